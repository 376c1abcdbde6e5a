// GENERATED: translation unit of solution 'test_partial_3d'.
#include "test_partial_3d.gen.cuh"
namespace yb { namespace gen { void test_partial_3d_register(GenStencil& g) { test_partial_3d_describe(g); } } }

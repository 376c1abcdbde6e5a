// GENERATED: translation unit of solution 'test_3d'.
#include "test_3d.gen.cuh"
namespace yb { namespace gen { void test_3d_register(GenStencil& g) { test_3d_describe(g); } } }

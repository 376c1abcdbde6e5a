// GENERATED: translation unit of solution 'test_stages_3d'.
#include "test_stages_3d.gen.cuh"
namespace yb { namespace gen { void test_stages_3d_register(GenStencil& g) { test_stages_3d_describe(g); } } }

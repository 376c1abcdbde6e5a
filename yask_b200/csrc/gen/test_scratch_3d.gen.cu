// GENERATED: translation unit of solution 'test_scratch_3d'.
#include "test_scratch_3d.gen.cuh"
namespace yb { namespace gen { void test_scratch_3d_register(GenStencil& g) { test_scratch_3d_describe(g); } } }

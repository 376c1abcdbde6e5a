// GENERATED: translation unit of solution 'test_scratch_2d'.
#include "test_scratch_2d.gen.cuh"
namespace yb { namespace gen { void test_scratch_2d_register(GenStencil& g) { test_scratch_2d_describe(g); } } }

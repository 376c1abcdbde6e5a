// GENERATED: translation unit of solution 'test_scratch_stages_1d'.
#include "test_scratch_stages_1d.gen.cuh"
namespace yb { namespace gen { void test_scratch_stages_1d_register(GenStencil& g) { test_scratch_stages_1d_describe(g); } } }

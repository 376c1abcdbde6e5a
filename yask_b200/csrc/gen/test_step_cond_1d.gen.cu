// GENERATED: translation unit of solution 'test_step_cond_1d'.
#include "test_step_cond_1d.gen.cuh"
namespace yb { namespace gen { void test_step_cond_1d_register(GenStencil& g) { test_step_cond_1d_describe(g); } } }

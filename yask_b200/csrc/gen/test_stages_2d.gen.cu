// GENERATED: translation unit of solution 'test_stages_2d'.
#include "test_stages_2d.gen.cuh"
namespace yb { namespace gen { void test_stages_2d_register(GenStencil& g) { test_stages_2d_describe(g); } } }

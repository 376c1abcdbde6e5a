// GENERATED: translation unit of solution 'test_2d'.
#include "test_2d.gen.cuh"
namespace yb { namespace gen { void test_2d_register(GenStencil& g) { test_2d_describe(g); } } }

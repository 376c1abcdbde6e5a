// GENERATED: translation unit of solution 'test_reverse_2d'.
#include "test_reverse_2d.gen.cuh"
namespace yb { namespace gen { void test_reverse_2d_register(GenStencil& g) { test_reverse_2d_describe(g); } } }

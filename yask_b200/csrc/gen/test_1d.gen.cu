// GENERATED: translation unit of solution 'test_1d'.
#include "test_1d.gen.cuh"
namespace yb { namespace gen { void test_1d_register(GenStencil& g) { test_1d_describe(g); } } }

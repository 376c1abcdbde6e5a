// GENERATED: translation unit of solution 'test_func_1d'.
#include "test_func_1d.gen.cuh"
namespace yb { namespace gen { void test_func_1d_register(GenStencil& g) { test_func_1d_describe(g); } } }

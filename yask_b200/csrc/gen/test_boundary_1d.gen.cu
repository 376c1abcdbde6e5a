// GENERATED: translation unit of solution 'test_boundary_1d'.
#include "test_boundary_1d.gen.cuh"
namespace yb { namespace gen { void test_boundary_1d_register(GenStencil& g) { test_boundary_1d_describe(g); } } }

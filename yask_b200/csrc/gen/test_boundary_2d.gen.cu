// GENERATED: translation unit of solution 'test_boundary_2d'.
#include "test_boundary_2d.gen.cuh"
namespace yb { namespace gen { void test_boundary_2d_register(GenStencil& g) { test_boundary_2d_describe(g); } } }

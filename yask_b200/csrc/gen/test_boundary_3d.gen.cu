// GENERATED: translation unit of solution 'test_boundary_3d'.
#include "test_boundary_3d.gen.cuh"
namespace yb { namespace gen { void test_boundary_3d_register(GenStencil& g) { test_boundary_3d_describe(g); } } }

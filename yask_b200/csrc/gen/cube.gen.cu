// GENERATED: translation unit of solution 'cube'.
#include "cube.gen.cuh"
namespace yb { namespace gen { void cube_register(GenStencil& g) { cube_describe(g); } } }

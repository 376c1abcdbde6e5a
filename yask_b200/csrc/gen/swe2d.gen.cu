// GENERATED: translation unit of solution 'swe2d'.
#include "swe2d.gen.cuh"
namespace yb { namespace gen { void swe2d_register(GenStencil& g) { swe2d_describe(g); } } }

// GENERATED: translation unit of solution 'wave2d'.
#include "wave2d.gen.cuh"
namespace yb { namespace gen { void wave2d_register(GenStencil& g) { wave2d_describe(g); } } }

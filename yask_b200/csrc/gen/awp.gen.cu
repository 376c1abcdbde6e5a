// GENERATED: translation unit of solution 'awp'.
#include "awp.gen.cuh"
namespace yb { namespace gen { void awp_register(GenStencil& g) { awp_describe(g); } } }

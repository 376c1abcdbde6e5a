// GENERATED: translation unit of solution 'awp_abc'.
#include "awp_abc.gen.cuh"
namespace yb { namespace gen { void awp_abc_register(GenStencil& g) { awp_abc_describe(g); } } }

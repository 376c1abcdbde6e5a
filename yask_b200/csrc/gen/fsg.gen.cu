// GENERATED: translation unit of solution 'fsg'.
#include "fsg.gen.cuh"
namespace yb { namespace gen { void fsg_register(GenStencil& g) { fsg_describe(g); } } }

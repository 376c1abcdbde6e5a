// GENERATED: translation unit of solution 'fsg2'.
#include "fsg2.gen.cuh"
namespace yb { namespace gen { void fsg2_register(GenStencil& g) { fsg2_describe(g); } } }

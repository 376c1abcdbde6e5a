// GENERATED: translation unit of solution 'fsg2_abc'.
#include "fsg2_abc.gen.cuh"
namespace yb { namespace gen { void fsg2_abc_register(GenStencil& g) { fsg2_abc_describe(g); } } }

// GENERATED: translation unit of solution 'fsg_abc'.
#include "fsg_abc.gen.cuh"
namespace yb { namespace gen { void fsg_abc_register(GenStencil& g) { fsg_abc_describe(g); } } }

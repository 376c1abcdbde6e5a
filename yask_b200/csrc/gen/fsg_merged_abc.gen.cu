// GENERATED: translation unit of solution 'fsg_merged_abc'.
#include "fsg_merged_abc.gen.cuh"
namespace yb { namespace gen { void fsg_merged_abc_register(GenStencil& g) { fsg_merged_abc_describe(g); } } }

// GENERATED: translation unit of solution 'fsg_merged'.
#include "fsg_merged.gen.cuh"
namespace yb { namespace gen { void fsg_merged_register(GenStencil& g) { fsg_merged_describe(g); } } }

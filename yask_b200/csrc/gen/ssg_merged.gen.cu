// GENERATED: translation unit of solution 'ssg_merged'.
#include "ssg_merged.gen.cuh"
namespace yb { namespace gen { void ssg_merged_register(GenStencil& g) { ssg_merged_describe(g); } } }

// GENERATED: translation unit of solution 'iso3dfd_fp64_r3'.
#include "iso3dfd_fp64_r3.gen.cuh"
namespace yb { namespace gen { void iso3dfd_fp64_r3_register(GenStencil& g) { iso3dfd_fp64_r3_describe(g); } } }

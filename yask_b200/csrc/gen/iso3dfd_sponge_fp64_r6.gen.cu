// GENERATED: translation unit of solution 'iso3dfd_sponge_fp64_r6'.
#include "iso3dfd_sponge_fp64_r6.gen.cuh"
namespace yb { namespace gen { void iso3dfd_sponge_fp64_r6_register(GenStencil& g) { iso3dfd_sponge_fp64_r6_describe(g); } } }

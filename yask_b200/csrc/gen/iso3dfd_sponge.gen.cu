// GENERATED: translation unit of solution 'iso3dfd_sponge'.
#include "iso3dfd_sponge.gen.cuh"
namespace yb { namespace gen { void iso3dfd_sponge_register(GenStencil& g) { iso3dfd_sponge_describe(g); } } }

// GENERATED: translation unit of solution 'iso3dfd_fp64'.
#include "iso3dfd_fp64.gen.cuh"
namespace yb { namespace gen { void iso3dfd_fp64_register(GenStencil& g) { iso3dfd_fp64_describe(g); } } }

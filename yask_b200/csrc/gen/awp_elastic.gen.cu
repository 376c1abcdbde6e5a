// GENERATED: translation unit of solution 'awp_elastic'.
#include "awp_elastic.gen.cuh"
namespace yb { namespace gen { void awp_elastic_register(GenStencil& g) { awp_elastic_describe(g); } } }

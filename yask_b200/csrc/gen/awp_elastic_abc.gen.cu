// GENERATED: translation unit of solution 'awp_elastic_abc'.
#include "awp_elastic_abc.gen.cuh"
namespace yb { namespace gen { void awp_elastic_abc_register(GenStencil& g) { awp_elastic_abc_describe(g); } } }

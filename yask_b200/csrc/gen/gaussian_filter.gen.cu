// GENERATED: translation unit of solution 'gaussian_filter'.
#include "gaussian_filter.gen.cuh"
namespace yb { namespace gen { void gaussian_filter_register(GenStencil& g) { gaussian_filter_describe(g); } } }

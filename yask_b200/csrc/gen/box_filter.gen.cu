// GENERATED: translation unit of solution 'box_filter'.
#include "box_filter.gen.cuh"
namespace yb { namespace gen { void box_filter_register(GenStencil& g) { box_filter_describe(g); } } }

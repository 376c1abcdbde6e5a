// GENERATED: translation unit of solution 'tti'.
#include "tti.gen.cuh"
namespace yb { namespace gen { void tti_register(GenStencil& g) { tti_describe(g); } } }

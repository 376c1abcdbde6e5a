// GENERATED: translation unit of solution 'ssg2'.
#include "ssg2.gen.cuh"
namespace yb { namespace gen { void ssg2_register(GenStencil& g) { ssg2_describe(g); } } }

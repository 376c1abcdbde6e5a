// GENERATED: translation unit of solution 'ssg'.
#include "ssg.gen.cuh"
namespace yb { namespace gen { void ssg_register(GenStencil& g) { ssg_describe(g); } } }

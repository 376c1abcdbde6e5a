// GENERATED: translation unit of solution '3plane'.
#include "s3plane.gen.cuh"
namespace yb { namespace gen { void s3plane_register(GenStencil& g) { s3plane_describe(g); } } }

// GENERATED: translation unit of solution '3axis'.
#include "s3axis.gen.cuh"
namespace yb { namespace gen { void s3axis_register(GenStencil& g) { s3axis_describe(g); } } }

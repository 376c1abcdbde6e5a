// GENERATED: translation unit of solution '3axis_with_diags'.
#include "s3axis_with_diags.gen.cuh"
namespace yb { namespace gen { void s3axis_with_diags_register(GenStencil& g) { s3axis_with_diags_describe(g); } } }

// GENERATED: translation unit of solution 'test_stream_3d'.
#include "test_stream_3d.gen.cuh"
namespace yb { namespace gen { void test_stream_3d_register(GenStencil& g) { test_stream_3d_describe(g); } } }

// GENERATED: translation unit of solution 'test_stream_2d'.
#include "test_stream_2d.gen.cuh"
namespace yb { namespace gen { void test_stream_2d_register(GenStencil& g) { test_stream_2d_describe(g); } } }

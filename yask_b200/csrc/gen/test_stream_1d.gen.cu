// GENERATED: translation unit of solution 'test_stream_1d'.
#include "test_stream_1d.gen.cuh"
namespace yb { namespace gen { void test_stream_1d_register(GenStencil& g) { test_stream_1d_describe(g); } } }

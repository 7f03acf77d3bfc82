// placeholder until the tcgen05 engine lands (next commit)
#include "common.cuh"
extern "C" int e4s_conv_tc(const E4SConv*, const void*, void*) { return e4s::fail(E4S_ERR_UNSUPPORTED, "conv_tc: not built"); }
extern "C" int64_t e4s_pack_weights_tc_bytes(int, int, int) { return 0; }
extern "C" int e4s_pack_weights_tc(const float*, int, int, int, int, void*, void*) { return e4s::fail(E4S_ERR_UNSUPPORTED, "conv_tc: not built"); }

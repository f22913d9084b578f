// Internal interface between api.cu and conv1x1.cu (separate translation units: they compile in parallel).
#pragma once
#include "../../include/hawq_b200.h"

namespace hawq {

// Try the stationary-weights 1x1 kernel (conv1x1.cuh).  Returns 0 = launched, 1 = not applicable (caller uses another kernel),
// HAWQ_ERR_CUDA on a tensor-map failure (message in c1_last_error()).
int launch_conv1x1(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w_ohwi,
                   const hawq_chan* chan, const void* res, void* out, void* out_low, int32_t* status, int sat_pack, void* stream);
int c1_set_attributes();
int c1_read_trace(long long* host_out, int n);   // debug timeline of the last conv1x1 launch (HAWQ_B200_HALO_TRACE=1)
const char* c1_last_error();

}  // namespace hawq

// Internal interface between api.cu and conv_halo.cu (separate translation units: they compile in parallel).
#pragma once
#include "../../include/hawq_b200.h"

namespace hawq {

// Try the in-place 3x3 kernel (conv_halo.cuh) for this launch.  Returns 0 = launched (2 = launched with the 2-D weight-map fallback), 1 = not applicable (caller uses
// another kernel), HAWQ_ERR_CUDA on a tensor-map / launch failure (message in halo_last_error()).
int launch_conv_halo(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w_ohwi,
                     const hawq_chan* chan, void* out, int32_t* status, void* stream);
int halo_set_attributes();
int halo_read_trace(long long* host_out, int n);   // debug timeline of the last conv_halo launch (HAWQ_B200_HALO_TRACE=1)
const char* halo_last_error();

}  // namespace hawq

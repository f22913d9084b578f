// Internal interface between api.cu and conv_dual.cu (separate translation units: they compile in parallel).
#pragma once
#include "../../include/hawq_b200.h"

namespace hawq {

// Try the stationary-weights dual-accumulator kernel (conv_dual.cuh) for a resize-unit tail.  Returns 0 = launched,
// 1 = not applicable (caller uses conv_tc), HAWQ_ERR_CUDA on a tensor-map failure (message in dual_last_error()).
int launch_conv_dual(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w,
                     const hawq_chan* chan, const hawq_conv_desc* d2, const void* x2, const int8_t* w2, const hawq_chan* chan2,
                     void* out, void* out_low, int32_t* status, int sat_pack, void* stream);
int dual_set_attributes();
const char* dual_last_error();

}  // namespace hawq

// Internal interface between api.cu and stem_tc.cu (separate translation units: they compile in parallel).
#pragma once
#include "../../include/hawq_b200.h"

namespace hawq {

// Fused tcgen05 stem (stem_tc.cuh).  Returns 0 = launched, 1 = not applicable, HAWQ_ERR_CUDA on a tensor-map failure.
int launch_stem_tc(int sm_count, int N, int H, int W, const int8_t* x, const int8_t* w256, const hawq_chan* chan, int clamp_lo, int clamp_hi,
                   int y_bits, void* y, int low_bits, uint32_t low_m, int low_e, int low_lo, int low_hi, void* out_low, int32_t* status,
                   void* stream);
int stem_tc_set_attributes();
const char* stem_tc_last_error();

}  // namespace hawq

// Host side of the fused tcgen05 stem (stem_tc.cuh): geometry, shared-memory plan, tensor maps, launch.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "stem_tc.cuh"
#include "stem_tc.h"

namespace hawq {

static thread_local char g_stem_err[256] = "";
const char* stem_tc_last_error() { return g_stem_err; }

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<encode_tiled_fn>(ptr);
  }();
  return fn;
}

constexpr int STEM_SMEM_MAX = 232448;
static int round_up(int v, int a) { return (v + a - 1) / a * a; }

int stem_tc_set_attributes() {
  const cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM_MAX);
  if (e != cudaSuccess) { snprintf(g_stem_err, sizeof(g_stem_err), "stem_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HAWQ_ERR_CUDA; }
  return HAWQ_OK;
}

#define ENC(map, dtype, rank, base, dims, strides, box, sw, what)                                                                          \
  do {                                                                                                                                    \
    const CUresult r_ = enc(&(map), dtype, rank, const_cast<void*>((const void*)(base)), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, \
                            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);                                    \
    if (r_ != CUDA_SUCCESS) { snprintf(g_stem_err, sizeof(g_stem_err), "stem_tc: cuTensorMapEncodeTiled (%s) failed (%d)", what, (int)r_); return HAWQ_ERR_CUDA; } \
  } while (0)

int launch_stem_tc(int sm_count, int N, int H, int W, const int8_t* x, const int8_t* w256, const hawq_chan* chan, int clamp_lo, int clamp_hi,
                   int y_bits, void* y, int low_bits, uint32_t low_m, int low_e, int low_lo, int low_hi, void* out_low, int32_t* status,
                   void* stream) {
  static const bool enabled = [] { const char* e = getenv("HAWQ_B200_STEMTC"); return !(e && e[0] == '0'); }();   // debugging switch
  if (!enabled) return 1;
  // rows are fetched as whole uint32 words by TMA: the row pitch 3 * W and the base address must be multiples of 16 bytes
  if (W % 16 != 0 || W > 256 || H < 8 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 1;
  if ((y_bits != 16 && y_bits != 32) || (low_bits != 0 && low_bits != 4 && low_bits != 8)) return 1;
  if (low_bits && low_m != 0u && (low_e < 31 || low_e > 51)) return 1;
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.chan = chan; p.status = status; p.N = N; p.H = H; p.W = W;
  p.Hc = (H + 6 - 7) / 2 + 1; p.Wc = (W + 6 - 7) / 2 + 1;
  p.Hp = (p.Hc + 2 - 3) / 2 + 1; p.Wp = (p.Wc + 2 - 3) / 2 + 1;
  if (p.Wc > 128 || p.Wp > 64) return 1;
  p.PB = 7; p.bands = (p.Hp + p.PB - 1) / p.PB;
  const long long units = (long long)N * p.bands;
  if (units > 0x7fffffff) return 1;
  p.units = (int)units;
  p.raw_rows = 4 * p.PB + 7; p.raw_pitch = W * 3; p.pix_pitch = W + 8;
  p.lo = clamp_lo; p.hi = clamp_hi; p.y_bits = y_bits;
  p.low_bits = low_bits; p.low_m = low_m; p.low_e = low_e; p.low_lo = low_lo; p.low_hi = low_hi;
  int total = 0;
  for (int obufs = 2; obufs >= 1; --obufs) {       // first choice: double-buffered output staging
    int off = 64 * 256;
    p.off_raw = off; off += round_up(p.raw_rows * p.raw_pitch, 1024);
    p.off_pix = off; off += round_up(p.raw_rows * p.pix_pitch * 4, 1024);
    p.off_a = off; off += 2 * STEM_A_TILE;
    p.off_rows = off; off += 3 * STEM_ROW_BYTES + round_up(p.Wc * 128, 1024);     // 4 row slots, the last one without the idle pixels
    p.y_stride = round_up(p.Wp * 64 * y_bits / 8, 1024); p.low_stride = round_up(p.Wp * 64, 1024); p.out_bufs = obufs;
    p.off_y = off; off += obufs * p.y_stride;
    p.off_low = off; off += obufs * p.low_stride;
    p.off_cst = off; off += 64 * 16;
    p.off_bar = off; off += 256;
    total = off + 1024;
    if (total <= STEM_SMEM_MAX) break;
  }
  if (total > STEM_SMEM_MAX) return 1;

  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) { snprintf(g_stem_err, sizeof(g_stem_err), "stem_tc: cuTensorMapEncodeTiled unavailable"); return HAWQ_ERR_CUDA; }
  StemMaps maps;
  memset(&maps, 0, sizeof(maps));
  const cuuint32_t ones[3] = {1, 1, 1};
  {
    const cuuint64_t words = (cuuint64_t)W * 3 / 4;
    const cuuint64_t dims[3] = {words, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t box[3] = {(cuuint32_t)words, (cuuint32_t)p.raw_rows, 1u};
    ENC(maps.x, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, "input");
  }
  {
    const cuuint64_t dims[3] = {64, 64, 4};
    const cuuint64_t strides[2] = {256, 64};
    const cuuint32_t box[3] = {64u, 64u, 4u};
    ENC(maps.w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, w256, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B, "weights");
  }
  const cuuint64_t rows = (cuuint64_t)N * p.Hp * p.Wp;
  if (y_bits == 16) {
    const cuuint64_t dims[2] = {128, rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {128u, (cuuint32_t)p.Wp};
    ENC(maps.y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "residual stream");
  } else {
    const cuuint64_t dims[3] = {128, rows, 2};
    const cuuint64_t strides[2] = {256, 128};
    const cuuint32_t box[3] = {128u, (cuuint32_t)p.Wp, 2u};
    ENC(maps.y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "residual stream (int32)");
  }
  if (low_bits) {
    const cuuint64_t rb = (cuuint64_t)64 * low_bits / 8;
    const cuuint64_t dims[2] = {rb, rows};
    const cuuint64_t strides[1] = {rb};
    const cuuint32_t box[2] = {(cuuint32_t)rb, (cuuint32_t)p.Wp};
    ENC(maps.low, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, out_low, dims, strides, box, rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, "low-bit output");
  }
  const int grid = (int)(units < sm_count ? units : sm_count);
  stem_tc_kernel<<<grid, STEM_TC_THREADS, total, (cudaStream_t)stream>>>(p, maps);
  return 0;
}

}  // namespace hawq

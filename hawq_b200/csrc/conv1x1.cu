// Host side of the stationary-weights 1x1 convolution (conv1x1.cuh): applicability, shared-memory plan, tensor maps, launch.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv1x1.cuh"
#include "conv1x1.h"

namespace hawq {

static thread_local char g_c1_err[256] = "";
static long long* g_c1_trace = nullptr;        // device buffer [4][48][4], allocated on first use when HAWQ_B200_HALO_TRACE=1

int c1_read_trace(long long* host_out, int n) {
  if (!g_c1_trace) return 0;
  if (n > 4 * 48 * 4) n = 4 * 48 * 4;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_c1_trace, sizeof(long long) * n, cudaMemcpyDeviceToHost);
  return n;
}
const char* c1_last_error() { return g_c1_err; }

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

constexpr int C1_SMEM_MAX = 232448;   // 227 KB

static int round_up(int v, int a) { return (v + a - 1) / a * a; }
static int largest_divisor_le(int n, int cap) {
  for (int d = cap; d > 1; --d)
    if (n % d == 0) return d;
  return 1;
}

struct C1Plan {
  int bn, total;
  C1Params p;
};

// shared-memory carve-up; prefers deep activation stages (fewer TMA operations), then more of them
static bool plan_for(int bn, bool a4, bool res, int KT, C1Plan* out) {
  C1Params& p = out->p;
  const int w_bytes = KT * bn * 64;
  const int res_bytes = res ? 2 * 128 * bn * 2 : 0;
  const int y_bytes = res ? 128 * bn * 2 : 0;      // staged uint16 output tile
  const int low_bytes = 128 * bn;                   // staged 8 / 4-bit output tile (REQUANT output, RESIDUAL low-bit copy)
  const int cst = bn * 16, bars = 256;
  for (int kc = 4; kc >= 1; --kc) {
    if (KT % kc) continue;
    const int a_stage = kc * 128 * 64;
    const int k_stage = a4 ? kc * 128 * 32 : 0;
    for (int ns = (kc == 4 ? 3 : C1_MAX_STAGES); ns >= 2; --ns) {
      int off = round_up(w_bytes, 1024);
      p.off_a = off; off += ns * a_stage;
      p.off_packed = off; off += ns * k_stage;
      off = round_up(off, 1024);
      p.off_res = off; off += res_bytes;
      p.off_y = off; off += y_bytes;
      p.off_low = off; off += low_bytes;
      p.off_cst = off; off += cst;
      p.off_bar = off; off += bars;
      const int total = off + 1024;
      if (total <= C1_SMEM_MAX) {
        out->bn = bn; out->total = total; p.NS = ns; p.KC = kc;
        p.w_box_kt = largest_divisor_le(KT, bn == 128 ? 8 : 16);
        p.w_boxes = KT / p.w_box_kt;
        return true;
      }
    }
  }
  return false;
}

template <int EPI, bool WIDE>
static cudaError_t set_attr() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(conv1x1_kernel<128, EPI, WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_MAX)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv1x1_kernel<64, EPI, WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_MAX)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(conv1x1_kernel<128, EPI, WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_MAX)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv1x1_kernel<64, EPI, WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_MAX);
}

int c1_set_attributes() {
  cudaError_t e;
  if ((e = set_attr<C1_REQ, false>()) != cudaSuccess || (e = set_attr<C1_RES, false>()) != cudaSuccess || (e = set_attr<C1_RES, true>()) != cudaSuccess) {
    snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return HAWQ_ERR_CUDA;
  }
  return HAWQ_OK;
}

template <int BN, int EPI, bool WIDE, bool A4>
static void launch(const C1Plan& plan, const C1Maps& maps, int grid, cudaStream_t st) {
  static const bool pdl = [] { const char* e = getenv("HAWQ_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(c1_threads(A4, EPI), 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, conv1x1_kernel<BN, EPI, WIDE, A4>, plan.p, maps);
}
template <int EPI, bool WIDE>
static void launch2(const C1Plan& plan, const C1Maps& maps, bool a4, int grid, cudaStream_t st) {
  if (plan.bn == 128) { if (a4) launch<128, EPI, WIDE, true>(plan, maps, grid, st); else launch<128, EPI, WIDE, false>(plan, maps, grid, st); }
  else { if (a4) launch<64, EPI, WIDE, true>(plan, maps, grid, st); else launch<64, EPI, WIDE, false>(plan, maps, grid, st); }
}

int launch_conv1x1(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w_ohwi,
                   const hawq_chan* chan, const void* res, void* out, void* out_low, int32_t* status, int sat_pack, void* stream) {
  static const bool enabled = [] { const char* e = getenv("HAWQ_B200_C1"); return !(e && e[0] == '0'); }();   // debugging switch
  if (!enabled) return 1;
  if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad != 0) return 1;
  const bool one = (ep->flags & HAWQ_EP_RATIOS_LE_ONE) != 0;
  const bool wide = !one && (ep->flags & HAWQ_EP_RATIOS_LE_2P20) != 0;
  bool is_res;
  if (ep->mode == HAWQ_EPI_REQUANT) {
    if ((ep->out_bits != 8 && ep->out_bits != 4) || !one) return 1;
    is_res = false;
  } else if (ep->mode == HAWQ_EPI_RESIDUAL) {
    if (ep->res_kind != 0 || ep->res_bits != 16 || ep->y_bits != 16 || !ep->relu || !(one || wide)) return 1;
    if (ep->res_m != 0u && ep->res_e > 51) return 1;
    if (ep->low_bits != 0 && (ep->low_e < 31 || ep->low_e > 51) && ep->low_m != 0u) return 1;
    is_res = true;
  } else {
    return 1;
  }
  const bool a4 = d->a_bits == 4;
  const int KT = d->Cin / 64;
  const long long M = (long long)d->N * d->H * d->W;
  if (M > 0x7fffff00ll) return 1;
  C1Plan plan;
  memset(&plan, 0, sizeof(plan));
  static const bool narrow = [] { const char* v = getenv("HAWQ_B200_BN"); return v && atoi(v) == 64; }();   // experiment switch: 64-channel blocks
  if (!((d->Cout % 128 == 0 && !narrow && plan_for(128, a4, is_res, KT, &plan)) || plan_for(64, a4, is_res, KT, &plan))) return 1;
  C1Params& p = plan.p;
  p.chan = chan; p.out = (uint8_t*)out; p.out_low = (uint8_t*)out_low; p.status = status;
  p.M = (int)M; p.Cout = d->Cout; p.KT = KT;
  p.m_tiles = (int)((M + 127) / 128); p.n_tiles = d->Cout / plan.bn;
  int per_n = sm_count / p.n_tiles;
  if (per_n < 1) return 1;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  p.ctas_per_n = per_n;
  p.relu = ep->relu; p.out_bits = ep->out_bits; p.lo = ep->clamp_lo; p.hi = ep->clamp_hi;
  p.res_m = ep->res_m; p.res_e = ep->res_e; p.low_bits = ep->low_bits; p.low_m = ep->low_m; p.low_e = ep->low_e;
  p.low_lo = ep->low_lo; p.low_hi = ep->low_hi; p.sat_pack = sat_pack;

  static const bool tracing = [] { const char* e = getenv("HAWQ_B200_HALO_TRACE"); return e && e[0] == '1'; }();
  if (tracing && !g_c1_trace) cudaMalloc(&g_c1_trace, sizeof(long long) * 4 * 48 * 4);
  if (tracing) cudaMemsetAsync(g_c1_trace, 0, sizeof(long long) * 4 * 48 * 4, (cudaStream_t)stream);
  p.trace = tracing ? g_c1_trace : nullptr;
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) { snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cuTensorMapEncodeTiled unavailable"); return HAWQ_ERR_CUDA; }
  C1Maps maps;
  memset(&maps, 0, sizeof(maps));
  const cuuint32_t estr[3] = {1, 1, 1};
  const uint64_t kt_bytes = a4 ? 32 : 64, row_bytes = (uint64_t)d->Cin * d->a_bits / 8;
  {
    const cuuint64_t dims[3] = {kt_bytes, (cuuint64_t)M, (cuuint64_t)KT};
    const cuuint64_t strides[2] = {row_bytes, kt_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)kt_bytes, 128u, (cuuint32_t)p.KC};
    const CUresult r = enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           a4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cuTensorMapEncodeTiled (activations) failed (%d)", (int)r); return HAWQ_ERR_CUDA; }
  }
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)d->Cout, (cuuint64_t)KT};
    const cuuint64_t strides[2] = {(cuuint64_t)d->Cin, 64};
    const cuuint32_t box[3] = {64u, (cuuint32_t)plan.bn, (cuuint32_t)p.w_box_kt};
    const CUresult r = enc(&maps.w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(w_ohwi), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cuTensorMapEncodeTiled (weights) failed (%d)", (int)r); return HAWQ_ERR_CUDA; }
  }
  if (is_res) {
    const cuuint64_t dims[3] = {128, (cuuint64_t)M, (cuuint64_t)d->Cout * 2 / 128};
    const cuuint64_t strides[2] = {(cuuint64_t)d->Cout * 2, 128};
    const cuuint32_t box[3] = {128u, 128u, (cuuint32_t)(plan.bn / 64)};
    CUresult r = enc(&maps.res, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(res), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = enc(&maps.y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cuTensorMapEncodeTiled (residual stream) failed (%d)", (int)r); return HAWQ_ERR_CUDA; }
  }
  {   // 8 / 4-bit output tile: REQUANT output, or the low-bit copy of the RESIDUAL epilogue
    const int bits = is_res ? ep->low_bits : ep->out_bits;
    void* base = is_res ? out_low : out;
    if (bits) {
      const cuuint32_t rb = (cuuint32_t)(plan.bn * bits / 8);        // 128 / 64 / 32
      const cuuint64_t dims[2] = {(cuuint64_t)d->Cout * bits / 8, (cuuint64_t)M};
      const cuuint64_t strides[1] = {(cuuint64_t)d->Cout * bits / 8};
      const cuuint32_t box[2] = {rb, 128u};
      const CUresult r = enc(&maps.low, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { snprintf(g_c1_err, sizeof(g_c1_err), "conv1x1: cuTensorMapEncodeTiled (low-bit output) failed (%d)", (int)r); return HAWQ_ERR_CUDA; }
    }
  }
  const int grid = p.n_tiles * p.ctas_per_n;
  cudaStream_t st = (cudaStream_t)stream;
  if (!is_res) launch2<C1_REQ, false>(plan, maps, a4, grid, st);
  else if (wide) launch2<C1_RES, true>(plan, maps, a4, grid, st);
  else launch2<C1_RES, false>(plan, maps, a4, grid, st);
  return 0;
}

}  // namespace hawq

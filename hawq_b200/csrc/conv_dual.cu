// Host side of the dual-accumulator resize-unit kernel (conv_dual.cuh): applicability, tile geometry, shared-memory plan, maps.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_dual.cuh"
#include "conv_dual.h"

namespace hawq {

static thread_local char g_dual_err[256] = "";
const char* dual_last_error() { return g_dual_err; }

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

constexpr int DUAL_SMEM_MAX = 232448;   // 227 KB

static int round_up(int v, int a) { return (v + a - 1) / a * a; }
static int largest_divisor_le(int n, int cap) {
  for (int d = cap; d > 1; --d)
    if (n % d == 0) return d;
  return 1;
}
static int gcd(int a, int b) { return b ? gcd(b, a % b) : a; }

struct DualPlan {
  int bn, total;
  DualParams p;
};

static bool plan_for(int bn, bool a4, int KT1, int KT2, int low_bits, DualPlan* out) {
  DualParams& p = out->p;
  const int w_bytes = (KT1 + KT2) * bn * 64;
  const int y_bytes = 128 * bn * 2, low_bytes = low_bits ? 128 * bn : 0;
  const int cst = 2 * bn * 16, bars = 256;
  const int g = gcd(KT1, KT2);
  for (int kc = 4; kc >= 1; --kc) {
    if (g % kc) continue;
    const int a_stage = kc * 128 * 64;
    for (int ns = (kc == 4 ? 3 : DUAL_MAX_STAGES); ns >= 2; --ns)
    for (int obufs = 2; obufs >= 1; --obufs) {       // double-buffered output staging where it costs no activation stage
      int off = round_up(w_bytes, 1024);
      p.off_a = off; off += ns * a_stage + 8192;        // + slack: the MMA reads 128 rows from k-tile blocks of TR rows
      p.off_packed = off; off += a4 ? ns * (a_stage / 2) : 0;
      p.off_y = off; off += obufs * y_bytes;
      p.off_low = off; off += obufs * low_bytes;
      p.out_bufs = obufs; p.y_stride = y_bytes; p.low_stride = low_bytes;
      p.off_cst = off; off += cst;
      p.off_bar = off; off += bars;
      const int total = off + 1024;
      if (total <= DUAL_SMEM_MAX) {
        out->bn = bn; out->total = total; p.NS = ns; p.KC = kc;
        p.w1_box_kt = largest_divisor_le(KT1, bn == 128 ? 8 : 16); p.w1_boxes = KT1 / p.w1_box_kt;
        p.w2_box_kt = largest_divisor_le(KT2, bn == 128 ? 8 : 16); p.w2_boxes = KT2 / p.w2_box_kt;
        return true;
      }
    }
  }
  return false;
}

int dual_set_attributes() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(conv_dual_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_dual_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DUAL_SMEM_MAX)) != cudaSuccess) {
    snprintf(g_dual_err, sizeof(g_dual_err), "conv_dual: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return HAWQ_ERR_CUDA;
  }
  return HAWQ_OK;
}

template <int BN, bool WIDE, bool A4>
static void launch(const DualPlan& plan, const DualMaps& maps, int grid, cudaStream_t st) {
  static const bool pdl = [] { const char* e = getenv("HAWQ_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(dual_threads(A4), 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, conv_dual_kernel<BN, WIDE, A4>, plan.p, maps);
}

#define ENC(map, rank, base, dims, strides, box, estr, sw, what)                                                                        \
  do {                                                                                                                                 \
    const CUresult r_ = enc(&(map), CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, const_cast<void*>((const void*)(base)), dims, strides, box, estr, \
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);     \
    if (r_ != CUDA_SUCCESS) {                                                                                                          \
      snprintf(g_dual_err, sizeof(g_dual_err), "conv_dual: cuTensorMapEncodeTiled (%s) failed (%d)", what, (int)r_);                   \
      return strict ? HAWQ_ERR_CUDA : 1;                                                                                               \
    }                                                                                                                                  \
  } while (0)

int launch_conv_dual(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w,
                     const hawq_chan* chan, const hawq_conv_desc* d2, const void* x2, const int8_t* w2, const hawq_chan* chan2,
                     void* out, void* out_low, int32_t* status, int sat_pack, void* stream) {
  static const bool enabled = [] { const char* e = getenv("HAWQ_B200_DUALK"); return !(e && e[0] == '0'); }();   // debugging switch
  if (!enabled) return 1;
  if ((d->a_bits != 8 && d->a_bits != 4) || d2->a_bits != d->a_bits) return 1;
  const bool a4 = d->a_bits == 4;
  const cuuint64_t ktb = a4 ? 32 : 64;                 // bytes of one k-tile (64 channels) in an activation row
  const bool one = (ep->flags & HAWQ_EP_RATIOS_LE_ONE) != 0;
  const bool wide = !one && (ep->flags & HAWQ_EP_RATIOS_LE_2P20) != 0;
  if (!one && !wide) return 1;
  if (ep->low_bits != 0 && ep->low_m != 0u && (ep->low_e < 31 || ep->low_e > 51)) return 1;
  const int Ho = d->H, Wo = d->W, s2 = d2->stride;
  const long long M = (long long)d->N * Ho * Wo;
  if (M > 0x7fffff00ll) return 1;
  if (s2 != 1 && (s2 != 2 || d2->H != 2 * Ho || d2->W != 2 * Wo)) return 1;
  // tile geometry: 128 linear rows (stride 1), or R whole output rows of one image / NI whole images (strided identity)
  int TR = 128, R = 0, NI = 1;
  if (s2 != 1) {
    if (Ho * Wo <= 128) { NI = 128 / (Ho * Wo); R = Ho; }
    else { R = largest_divisor_le(Ho, 128 / Wo); if (128 / Wo < 1) return 1; }
    TR = R * Wo * NI;
    if (TR < 64 || 2 * Wo > 256 || 2 * R > 256) return 1;
  }
  const int KT1 = d->Cin / 64, KT2 = d2->Cin / 64;
  DualPlan plan;
  memset(&plan, 0, sizeof(plan));
  if (!((d->Cout % 128 == 0 && plan_for(128, a4, KT1, KT2, ep->low_bits, &plan)) || plan_for(64, a4, KT1, KT2, ep->low_bits, &plan))) return 1;
  DualParams& p = plan.p;
  p.chan = chan; p.chan2 = chan2; p.status = status;
  p.M = (int)M; p.Cout = d->Cout; p.KT1 = KT1; p.KT2 = KT2; p.TR = TR;
  p.strided = s2 != 1; p.Wo = Wo; p.HoWo = Ho * Wo; p.R = R; p.stride2 = s2;
  p.m_tiles = (int)((M + TR - 1) / TR); p.n_tiles = d->Cout / plan.bn;
  int per_n = sm_count / p.n_tiles;
  if (per_n < 1) return 1;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  p.ctas_per_n = per_n;
  p.low_bits = ep->low_bits; p.low_m = ep->low_m; p.low_e = ep->low_e; p.low_lo = ep->low_lo; p.low_hi = ep->low_hi; p.sat_pack = sat_pack;

  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) { snprintf(g_dual_err, sizeof(g_dual_err), "conv_dual: cuTensorMapEncodeTiled unavailable"); return HAWQ_ERR_CUDA; }
  DualMaps maps;
  memset(&maps, 0, sizeof(maps));
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  bool strict = true;
  const CUtensorMapSwizzle sw_a = a4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    const cuuint64_t dims[3] = {ktb, (cuuint64_t)M, (cuuint64_t)KT1};
    const cuuint64_t strides[2] = {(cuuint64_t)d->Cin * d->a_bits / 8, ktb};
    const cuuint32_t box[3] = {(cuuint32_t)ktb, (cuuint32_t)TR, (cuuint32_t)p.KC};
    ENC(maps.a, 3, x, dims, strides, box, ones, sw_a, "activations");
  }
  if (s2 == 1) {
    const cuuint64_t dims[3] = {ktb, (cuuint64_t)M, (cuuint64_t)KT2};
    const cuuint64_t strides[2] = {(cuuint64_t)d2->Cin * d->a_bits / 8, ktb};
    const cuuint32_t box[3] = {(cuuint32_t)ktb, (cuuint32_t)TR, (cuuint32_t)p.KC};
    ENC(maps.a2, 3, x2, dims, strides, box, ones, sw_a, "identity activations");
  } else {
    strict = false;           // a driver that refuses the strided 5-D view: the caller falls back to conv_tc
    const cuuint64_t c2 = (cuuint64_t)d2->Cin * d->a_bits / 8;
    const cuuint64_t dims[5] = {ktb, (cuuint64_t)d2->W, (cuuint64_t)d2->H, (cuuint64_t)d2->N, (cuuint64_t)KT2};
    const cuuint64_t strides[4] = {c2, c2 * d2->W, c2 * d2->W * d2->H, ktb};
    const cuuint32_t box[5] = {(cuuint32_t)ktb, (cuuint32_t)(2 * Wo), (cuuint32_t)(2 * R), (cuuint32_t)NI, (cuuint32_t)p.KC};
    const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
    ENC(maps.a2, 5, x2, dims, strides, box, estr, sw_a, "strided identity activations");
    strict = true;
  }
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)d->Cout, (cuuint64_t)KT1};
    const cuuint64_t strides[2] = {(cuuint64_t)d->Cin, 64};
    const cuuint32_t box[3] = {64u, (cuuint32_t)plan.bn, (cuuint32_t)p.w1_box_kt};
    ENC(maps.w1, 3, w, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_64B, "weights");
  }
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)d->Cout, (cuuint64_t)KT2};
    const cuuint64_t strides[2] = {(cuuint64_t)d2->Cin, 64};
    const cuuint32_t box[3] = {64u, (cuuint32_t)plan.bn, (cuuint32_t)p.w2_box_kt};
    ENC(maps.w2, 3, w2, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_64B, "identity weights");
  }
  {
    const cuuint64_t dims[3] = {128, (cuuint64_t)M, (cuuint64_t)d->Cout * 2 / 128};
    const cuuint64_t strides[2] = {(cuuint64_t)d->Cout * 2, 128};
    const cuuint32_t box[3] = {128u, (cuuint32_t)TR, (cuuint32_t)(plan.bn / 64)};
    ENC(maps.y, 3, out, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B, "residual stream");
  }
  if (ep->low_bits) {
    const cuuint32_t rb = (cuuint32_t)(plan.bn * ep->low_bits / 8);
    const cuuint64_t dims[2] = {(cuuint64_t)d->Cout * ep->low_bits / 8, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)d->Cout * ep->low_bits / 8};
    const cuuint32_t box[2] = {rb, (cuuint32_t)TR};
    ENC(maps.low, 2, out_low, dims, strides, box, ones, rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
        "low-bit output");
  }
  const int grid = p.n_tiles * p.ctas_per_n;
  cudaStream_t st = (cudaStream_t)stream;
  if (a4) {
    if (plan.bn == 128) { if (wide) launch<128, true, true>(plan, maps, grid, st); else launch<128, false, true>(plan, maps, grid, st); }
    else { if (wide) launch<64, true, true>(plan, maps, grid, st); else launch<64, false, true>(plan, maps, grid, st); }
  } else {
    if (plan.bn == 128) { if (wide) launch<128, true, false>(plan, maps, grid, st); else launch<128, false, false>(plan, maps, grid, st); }
    else { if (wide) launch<64, true, false>(plan, maps, grid, st); else launch<64, false, false>(plan, maps, grid, st); }
  }
  return 0;
}

}  // namespace hawq

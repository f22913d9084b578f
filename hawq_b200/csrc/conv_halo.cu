// Host side of the in-place 3x3 convolution (conv_halo.cuh): applicability, shared-memory plan, 4-D tensor map, launch.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_halo.cuh"
#include "conv_halo.h"

namespace hawq {

static thread_local char g_halo_err[256] = "";
static long long* g_halo_trace = nullptr;      // device buffer [4][64][4], allocated on first use when HAWQ_B200_HALO_TRACE=1

int halo_read_trace(long long* host_out, int n) {
  if (!g_halo_trace) return 0;
  if (n > 4 * 64 * 4) n = 4 * 64 * 4;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_halo_trace, sizeof(long long) * n, cudaMemcpyDeviceToHost);
  return n;
}
const char* halo_last_error() { return g_halo_err; }

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

constexpr int HALO_SMEM_MAX = 232448;   // 227 KB

struct HaloPlan {
  int bn, npb, nkb, total;
  HaloParams p;
};

static int round_up(int v, int a) { return (v + a - 1) / a * a; }

// shared-memory carve-up for channel-block width bn; returns false if it does not fit
static bool plan_for(int bn, bool a4, int K, int wp, int R, HaloPlan* out) {
  const int b_bytes = (K / 64) * bn * 64;
  const int patch_alloc = round_up((128 + 2 * wp + 2) * 64, 1024);
  const int packed_alloc = a4 ? round_up((R + 2) * wp * 32, 1024) : 0;
  const int out_tile = 128 * bn;       // staged output tile (8-bit worst case)
  const int cst = bn * 16;
  const int bars = 256;
  for (int npb = a4 ? 2 : HALO_MAX_BUFS; npb >= 2; --npb) {
    const int nkb = a4 ? 3 : 0;
    int off = round_up(b_bytes, 1024);
    HaloParams& p = out->p;
    p.off_patch = off; off += npb * patch_alloc;
    p.off_packed = off; off += nkb * packed_alloc;
    off = round_up(off, 1024);
    p.off_out = off; off += out_tile;
    p.off_cst = off; off += cst;
    p.off_bar = off; off += bars;
    const int total = off + 1024;      // slack for the 1024-byte alignment of the base
    if (total <= HALO_SMEM_MAX) {
      out->bn = bn; out->npb = npb; out->nkb = nkb; out->total = total;
      p.patch_alloc = patch_alloc; p.packed_alloc = packed_alloc; p.npb = npb; p.nkb = nkb;
      return true;
    }
  }
  return false;
}

int halo_set_attributes() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(conv_halo_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_halo_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX)) != cudaSuccess) {
    snprintf(g_halo_err, sizeof(g_halo_err), "conv_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return HAWQ_ERR_CUDA;
  }
  return HAWQ_OK;
}

template <int BN, bool A4>
static void launch(const HaloPlan& plan, const CUtensorMap& map, const CUtensorMap& wmap, const CUtensorMap& omap, int grid, cudaStream_t st) {
  static const bool pdl = [] { const char* e = getenv("HAWQ_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(halo_threads(A4), 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (plan.p.trace) cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, A4, true>, plan.p, map, wmap, omap);
  else cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, A4, false>, plan.p, map, wmap, omap);
}

int launch_conv_halo(int sm_count, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w_ohwi,
                     const hawq_chan* chan, void* out, int32_t* status, void* stream) {
  static const bool enabled = [] { const char* e = getenv("HAWQ_B200_HALO"); return !(e && e[0] == '0'); }();   // debugging switch
  if (!enabled) return 1;
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) return 1;
  if (ep->mode != HAWQ_EPI_REQUANT || (ep->out_bits != 8 && ep->out_bits != 4) || !(ep->flags & HAWQ_EP_RATIOS_LE_ONE)) return 1;
  const int wp = d->W + 2;
  if (wp > 128 || wp > 256) return 1;
  const bool a4 = d->a_bits == 4;
  const int R = d->H < 128 / wp ? d->H : 128 / wp;
  const int K = 9 * d->Cin;
  HaloPlan plan;
  memset(&plan, 0, sizeof(plan));
  static const bool narrow = [] { const char* v = getenv("HAWQ_B200_BN"); return v && atoi(v) == 64; }();   // experiment switch: 64-channel blocks
  if (!((d->Cout % 128 == 0 && !narrow && plan_for(128, a4, K, wp, R, &plan)) || plan_for(64, a4, K, wp, R, &plan))) return 1;
  HaloParams& p = plan.p;
  p.chan = chan; p.out = (uint8_t*)out; p.status = status;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cout = d->Cout; p.chunks = d->Cin / 64; p.R = R; p.wp = wp;
  p.tiles_per_img = (d->H + R - 1) / R;
  const long long m_tiles = (long long)d->N * p.tiles_per_img;
  if (m_tiles > 0x7fffffff) return 1;
  p.m_tiles = (int)m_tiles; p.n_tiles = d->Cout / plan.bn;
  int per_n = sm_count / p.n_tiles;
  if (per_n < 1) return 1;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  p.ctas_per_n = per_n;
  p.patch_bytes = (R + 2) * wp * (a4 ? 32 : 64);
  p.relu = ep->relu; p.out_bits = ep->out_bits; p.lo = ep->clamp_lo; p.hi = ep->clamp_hi;

  static const bool tracing = [] { const char* e = getenv("HAWQ_B200_HALO_TRACE"); return e && e[0] == '1'; }();
  if (tracing && !g_halo_trace) { cudaMalloc(&g_halo_trace, sizeof(long long) * 4 * 64 * 4); }
  if (tracing) cudaMemsetAsync(g_halo_trace, 0, sizeof(long long) * 4 * 64 * 4, (cudaStream_t)stream);
  p.trace = tracing ? g_halo_trace : nullptr;
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) { snprintf(g_halo_err, sizeof(g_halo_err), "conv_halo: cuTensorMapEncodeTiled unavailable"); return HAWQ_ERR_CUDA; }
  CUtensorMap map;
  const uint64_t cb = (uint64_t)d->Cin * d->a_bits / 8;       // bytes per pixel
  const cuuint64_t dims[4] = {cb, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
  const cuuint64_t strides[3] = {cb, cb * d->W, cb * d->W * d->H};
  const cuuint32_t box[4] = {a4 ? 32u : 64u, (cuuint32_t)wp, (cuuint32_t)(R + 2), 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         a4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_halo_err, sizeof(g_halo_err), "conv_halo: cuTensorMapEncodeTiled failed (%d)", (int)r); return HAWQ_ERR_CUDA; }

  // weights [Cout][3][3][Cin] int8 as {Cin bytes, Cout rows (pitch 9 * Cin), 9 taps (pitch Cin)}: box {64, BN, 9}
  CUtensorMap wmap;
  const cuuint64_t wdims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Cout, 9};
  const cuuint64_t wstrides[2] = {(cuuint64_t)K, (cuuint64_t)d->Cin};
  const cuuint32_t wbox[3] = {64u, (cuuint32_t)plan.bn, 9u};
  const cuuint32_t westr[3] = {1, 1, 1};
  const CUresult rw = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(w_ohwi), wdims, wstrides, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  p.w_rank3 = 1;
  if (rw != CUDA_SUCCESS) {        // plain matrix view [Cout][K], box {64, BN}
    const cuuint64_t mdims[2] = {(cuuint64_t)K, (cuuint64_t)d->Cout};
    const cuuint64_t mstrides[1] = {(cuuint64_t)K};
    const cuuint32_t mbox[2] = {64u, (cuuint32_t)plan.bn};
    const CUresult r2 = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(w_ohwi), mdims, mstrides, mbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r2 != CUDA_SUCCESS) { snprintf(g_halo_err, sizeof(g_halo_err), "conv_halo: cuTensorMapEncodeTiled (weights) failed (%d, %d)", (int)rw, (int)r2); return HAWQ_ERR_CUDA; }
    p.w_rank3 = 0;
  }

  // output [N][H][W][Cout * bits / 8] as {bytes, W, H, N}: store box {BN * bits / 8, W + 2, R, 1} (x >= W and y >= H are clipped)
  CUtensorMap omap;
  {
    const cuuint64_t ob = (cuuint64_t)d->Cout * ep->out_bits / 8;
    const cuuint32_t rb = (cuuint32_t)(plan.bn * ep->out_bits / 8);
    const cuuint64_t odims[4] = {ob, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    const cuuint64_t ostrides[3] = {ob, ob * d->W, ob * d->W * d->H};
    const cuuint32_t obox[4] = {rb, (cuuint32_t)wp, (cuuint32_t)R, 1u};
    const CUresult ro = enc(&omap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, out, odims, ostrides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (ro != CUDA_SUCCESS) { snprintf(g_halo_err, sizeof(g_halo_err), "conv_halo: cuTensorMapEncodeTiled (output) failed (%d)", (int)ro); return HAWQ_ERR_CUDA; }
  }
  const int grid = p.n_tiles * p.ctas_per_n;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan.bn == 128) { if (a4) launch<128, true>(plan, map, wmap, omap, grid, st); else launch<128, false>(plan, map, wmap, omap, grid, st); }
  else { if (a4) launch<64, true>(plan, map, wmap, omap, grid, st); else launch<64, false>(plan, map, wmap, omap, grid, st); }
  return p.w_rank3 ? 0 : 2;
}

}  // namespace hawq

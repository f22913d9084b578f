// C ABI of libhawq_b200.so (see include/hawq_b200.h).  Argument validation + kernel launches; no allocation,
// no synchronisation on the data path (CUDA-graph safe).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "conv1x1.h"
#include "conv_dual.h"
#include "conv_halo.h"
#include "conv_igemm.cuh"
#include "conv_tc.cuh"
#include "elementwise.cuh"
#include "stem.cuh"
#include "stem_tc.h"

using namespace hawq;

struct hawq_handle {
  int device;
  int sm_count;
  int32_t* status;  // device status word (owned)
};

static thread_local char g_err[512] = "";
static long long g_kernel_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // hawq_debug_kernel_count: launches by kernel family (not atomic: debugging aid)

// ---- TMA tensor maps (driver entry point resolved through the runtime: no link-time dependency on libcuda)
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
// byte matrix [rows][inner_bytes] with row pitch pitch_bytes; box = box_inner bytes x box_rows rows
static int make_map_2d(CUtensorMap* m, const void* base, uint64_t inner_bytes, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner,
                       uint32_t box_rows, CUtensorMapSwizzle sw) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return -1;
  const cuuint64_t dims[2] = {inner_bytes, rows};
  const cuuint64_t strides[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return fail(HAWQ_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HAWQ_ERR_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
  return HAWQ_OK;
}

static int grid_for(long long work_items, int sm_count) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int BN, bool A4, int EPI>
static int set_conv_attr1() {
  CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<BN, A4, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ConvSmem<BN, A4>::TOTAL));
  return HAWQ_OK;
}
template <int BN, bool A4>
static int set_conv_attr() {
  int rc;
  if ((rc = set_conv_attr1<BN, A4, EPI_GENERIC>()) || (rc = set_conv_attr1<BN, A4, EPI_FAST_LOW>()) ||
      (rc = set_conv_attr1<BN, A4, EPI_FAST_RES>()))
    return rc;
  return HAWQ_OK;
}

template <int EPI, bool WIDE>
static int set_tc_attr1() {
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<128, EPI, WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128, EPI, false>::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<64, EPI, WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64, EPI, false>::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<128, EPI, WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128, EPI, true>::TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<64, EPI, WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64, EPI, true>::TOTAL));
  return HAWQ_OK;
}
template <int EPI>
static int set_tc_attr() {
  int rc = set_tc_attr1<EPI, false>();
  if (rc == HAWQ_OK && EPI >= TC_EPI_RES22) rc = set_tc_attr1<EPI, true>();
  return rc;
}

// conv_tc launches carry the programmatic-stream-serialization attribute (HAWQ_B200_PDL != 0): the kernel's prologue may
// start while the previous kernel of the stream drains; the kernel itself waits (griddepcontrol.wait) before touching memory
template <int BN, int EPI, bool WIDE, bool A4>
static void launch_tc3(const ConvParams& p, const TcMaps& maps, int grid, cudaStream_t st) {
  static const bool pdl = [] { const char* e = getenv("HAWQ_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = TcSmem<BN, EPI, A4>::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, EPI, WIDE, A4>, p, maps);   // errors surface through launch_check()
}
template <int EPI, bool WIDE, bool A4>
static void launch_tc2(const ConvParams& p, const TcMaps& maps, bool bn128, int grid, cudaStream_t st) {
  if (bn128) launch_tc3<128, EPI, WIDE, A4>(p, maps, grid, st);
  else launch_tc3<64, EPI, WIDE, A4>(p, maps, grid, st);
}
template <int EPI>
static void launch_tc(const ConvParams& p, const TcMaps& maps, bool bn128, bool ratios_wide, bool a4, int grid, cudaStream_t st) {
  if constexpr (EPI >= TC_EPI_RES22) {
    if (ratios_wide) {
      if (a4) launch_tc2<EPI, true, true>(p, maps, bn128, grid, st);
      else launch_tc2<EPI, true, false>(p, maps, bn128, grid, st);
      return;
    }
  }
  if (a4) launch_tc2<EPI, false, true>(p, maps, bn128, grid, st);
  else launch_tc2<EPI, false, false>(p, maps, bn128, grid, st);
}

template <int BN, bool A4>
static void launch_conv(const ConvParams& p, dim3 grid, cudaStream_t s) {
  const int smem = ConvSmem<BN, A4>::TOTAL;
  if (p.mode == HAWQ_EPI_REQUANT && p.out_bits <= 8) conv_igemm_kernel<BN, A4, EPI_FAST_LOW><<<grid, CONV_THREADS, smem, s>>>(p);
  else if (p.mode == HAWQ_EPI_RESIDUAL) conv_igemm_kernel<BN, A4, EPI_FAST_RES><<<grid, CONV_THREADS, smem, s>>>(p);
  else conv_igemm_kernel<BN, A4, EPI_GENERIC><<<grid, CONV_THREADS, smem, s>>>(p);
}

extern "C" {

int hawq_abi_version(void) { return HAWQ_ABI_VERSION; }
const char* hawq_last_error(void) { return g_err; }

int hawq_create(int device, hawq_handle** out) {
  if (!out) return fail(HAWQ_ERR_BAD_ARG, "hawq_create: out is null");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_create: device sm_%d%d is not Blackwell sm_100", prop.major, prop.minor);
  hawq_handle* h = new hawq_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->status = nullptr;
  CUDA_TRY(cudaMalloc(&h->status, sizeof(int32_t)));
  CUDA_TRY(cudaMemset(h->status, 0, sizeof(int32_t)));
  int rc;
  if ((rc = set_tc_attr<TC_EPI_REQ>()) || (rc = set_tc_attr<TC_EPI_RAW>()) || (rc = set_tc_attr<TC_EPI_RES22>()) ||
      (rc = set_tc_attr<TC_EPI_RES44>()) || (rc = set_tc_attr<TC_EPI_RES42>()) || (rc = set_tc_attr<TC_EPI_DUAL>()))
    return rc;
  if ((rc = halo_set_attributes())) return fail(rc, "%s", halo_last_error());
  if ((rc = c1_set_attributes())) return fail(rc, "%s", c1_last_error());
  if ((rc = dual_set_attributes())) return fail(rc, "%s", dual_last_error());
  if ((rc = stem_tc_set_attributes())) return fail(rc, "%s", stem_tc_last_error());
  CUDA_TRY(cudaFuncSetAttribute(linear_dp4a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, linear_smem_bytes(LIN_MAX_K)));
  if ((rc = set_conv_attr<128, false>()) || (rc = set_conv_attr<64, false>()) || (rc = set_conv_attr<128, true>()) ||
      (rc = set_conv_attr<64, true>()))
    return rc;
  *out = h;
  return HAWQ_OK;
}

int hawq_destroy(hawq_handle* h) {
  if (!h) return HAWQ_OK;
  cudaFree(h->status);
  delete h;
  return HAWQ_OK;
}

int hawq_sm_count(const hawq_handle* h) { return h ? h->sm_count : 0; }

int hawq_reset_status(hawq_handle* h, void* stream) {
  if (!h) return fail(HAWQ_ERR_BAD_ARG, "null handle");
  CUDA_TRY(cudaMemsetAsync(h->status, 0, sizeof(int32_t), (cudaStream_t)stream));
  return HAWQ_OK;
}

int hawq_get_status(hawq_handle* h, void* stream, int32_t* host_flags) {
  if (!h || !host_flags) return fail(HAWQ_ERR_BAD_ARG, "null argument");
  CUDA_TRY(cudaMemcpyAsync(host_flags, h->status, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return HAWQ_OK;
}

int hawq_copy_status(hawq_handle* h, int32_t* dst, void* stream) {
  if (!h || !dst) return fail(HAWQ_ERR_BAD_ARG, "null argument");
  CUDA_TRY(cudaMemcpyAsync(dst, h->status, sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HAWQ_OK;
}

static bool sat_pack_enabled() {
  static const bool on = [] { const char* e = getenv("HAWQ_B200_SATPACK"); return !(e && e[0] == '0'); }();
  return on;
}

// scalar dyadic pairs whose folded FMA constant (magic - 2^52 * m * 2^-e) is exact in the tcgen05 RESIDUAL epilogues
static bool fold_ok(uint32_t m, int e) { return m == 0u || e <= 51; }

static int check_me(uint32_t m, int e, const char* what) {
  if (e < 1 || e > 62 || m > 0x80000000u) return fail(HAWQ_ERR_BAD_ARG, "%s: dyadic pair out of range (m=%u e=%d)", what, m, e);
  return HAWQ_OK;
}

int hawq_conv2d(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w,
                const hawq_chan* chan, const void* res, const hawq_chan* res_chan, const float* fscale, void* out,
                void* out_low, void* stream) {
  if (!h || !d || !ep || !x || !w || !chan) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: null argument");
  if (d->N < 1 || d->H < 1 || d->W < 1 || d->kh < 1 || d->kw < 1 || d->stride < 1 || d->pad < 0)
    return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: bad geometry");
  if (d->a_bits != 8 && d->a_bits != 4) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d: a_bits must be 4 or 8");
  if (d->Cin % 64 != 0 || d->Cout % 64 != 0)
    return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d: Cin (%d) and Cout (%d) must be multiples of 64", d->Cin, d->Cout);
  const int Ho = (d->H + 2 * d->pad - d->kh) / d->stride + 1;
  const int Wo = (d->W + 2 * d->pad - d->kw) / d->stride + 1;
  if (Ho < 1 || Wo < 1) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: empty output");
  const long long M = (long long)d->N * Ho * Wo;
  if (M > 0x7fffff00ll || (long long)d->N * d->H * d->W > 0x7fffff00ll) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d: too many pixels");

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const uint8_t*)x; p.w = w; p.chan = chan; p.res = res; p.res_chan = res_chan; p.fscale = fscale;
  p.out = out; p.out_low = out_low; p.status = h->status;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.KH = d->kh; p.KW = d->kw;
  p.stride = d->stride; p.pad = d->pad; p.Ho = Ho; p.Wo = Wo; p.M = (int)M; p.K = d->kh * d->kw * d->Cin;
  p.cin_chunks = d->Cin / 64;
  p.x_pix_bytes = d->Cin * d->a_bits / 8;
  p.mode = ep->mode; p.relu = ep->relu; p.out_bits = ep->out_bits; p.lo = ep->clamp_lo; p.hi = ep->clamp_hi;
  p.res_kind = ep->res_kind; p.res_bits = ep->res_bits; p.res_m = ep->res_m; p.res_e = ep->res_e;
  p.y_bits = ep->y_bits; p.low_bits = ep->low_bits; p.low_m = ep->low_m; p.low_e = ep->low_e;
  p.low_lo = ep->low_lo; p.low_hi = ep->low_hi; p.cout_store = ep->cout_store;
  p.slow_scalar = 0;
  p.tma_a = 0;
  p.tma_io = 0;
  p.patch_rows = 0;
  p.w_tiled = nullptr;
  p.sat_pack = sat_pack_enabled() ? 1 : 0;
  if (ep->mode == HAWQ_EPI_RESIDUAL) {
    if (ep->res_kind == 0 && !dyadic_is_fast(ep->res_m, ep->res_e)) p.slow_scalar = 1;
    if (ep->low_bits != 0 && !dyadic_is_fast(ep->low_m, ep->low_e)) p.slow_scalar = 1;
  }

  switch (ep->mode) {
    case HAWQ_EPI_REQUANT:
      if (!out) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: REQUANT needs out");
      if (ep->out_bits != 4 && ep->out_bits != 8 && ep->out_bits != 16 && ep->out_bits != 32)
        return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: out_bits must be 4/8/16/32");
      if (ep->clamp_lo > ep->clamp_hi) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: empty clamp range");
      break;
    case HAWQ_EPI_RESIDUAL:
      if (!res) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: RESIDUAL needs res");
      if (ep->res_kind == 1) {
        if (!res_chan) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: res_kind 1 needs res_chan");
      } else if (ep->res_kind == 0) {
        if (ep->res_bits != 16 && ep->res_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: res_bits must be 16/32");
        int rc = check_me(ep->res_m, ep->res_e, "hawq_conv2d residual");
        if (rc) return rc;
      } else return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: res_kind must be 0/1");
      if (ep->y_bits != 0 && ep->y_bits != 16 && ep->y_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: y_bits must be 0/16/32");
      if (ep->y_bits == 16 && !ep->relu) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: uint16 residual stream requires relu");
      if (ep->y_bits && !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: y_bits set but out is null");
      if (ep->low_bits != 0 && ep->low_bits != 4 && ep->low_bits != 8) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: low_bits must be 0/4/8");
      if (ep->low_bits) {
        if (!out_low) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: low_bits set but out_low is null");
        int rc = check_me(ep->low_m, ep->low_e, "hawq_conv2d low-bit copy");
        if (rc) return rc;
      }
      if (!ep->y_bits && !ep->low_bits) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: RESIDUAL with no output");
      break;
    case HAWQ_EPI_RAW_I32:
      if (!out) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: RAW_I32 needs out");
      break;
    case HAWQ_EPI_DEQUANT_F32:
      if (!out || !fscale) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: DEQUANT_F32 needs out and fscale");
      if (ep->cout_store < 1 || ep->cout_store > d->Cout) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: bad cout_store");
      break;
    default:
      return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d: unknown epilogue mode %d", ep->mode);
  }

  // tcgen05 path: int8 activations, the three hot epilogues, all ratios <= 1 (promised), HAWQ_B200_TC != 0
  static const bool tc_enabled = [] { const char* e = getenv("HAWQ_B200_TC"); return !(e && e[0] == '0'); }();
  const bool tc_epi = (ep->mode == HAWQ_EPI_REQUANT && ep->out_bits <= 8) || (ep->mode == HAWQ_EPI_RESIDUAL && ep->y_bits != 0 && ep->relu) ||
                      ep->mode == HAWQ_EPI_RAW_I32;
  const bool ratios_one = (ep->flags & HAWQ_EP_RATIOS_LE_ONE) != 0;
  const bool ratios_wide = !ratios_one && (ep->flags & HAWQ_EP_RATIOS_LE_2P20) != 0 && ep->mode == HAWQ_EPI_RESIDUAL;
  const bool folds = ep->mode != HAWQ_EPI_RESIDUAL || ((ep->res_kind != 0 || fold_ok(ep->res_m, ep->res_e)) && (ep->low_bits == 0 || fold_ok(ep->low_m, ep->low_e)));
  if (tc_enabled && tc_epi && folds && (ratios_one || ratios_wide)) {
    const bool a4 = d->a_bits == 4;
    // 3x3 stride-1 REQUANT layers: A operand read in place from a zero-padded TMA patch, weights stationary (conv_halo.cuh)
    if (ep->mode == HAWQ_EPI_REQUANT) {
      const int hr = launch_conv_halo(h->sm_count, d, ep, x, w, chan, out, h->status, stream);
      if (hr < 0) return fail(hr, "%s", halo_last_error());
      if (hr == 0 || hr == 2) { ++g_kernel_count[1]; g_kernel_count[2] += hr == 2; return launch_check("conv_halo"); }
    }
    // 1x1 stride-1 layers (REQUANT, uint16-stream RESIDUAL): stationary weights, large TMA copies (conv1x1.cuh)
    if (ep->mode == HAWQ_EPI_REQUANT || ep->mode == HAWQ_EPI_RESIDUAL) {
      const int cr = launch_conv1x1(h->sm_count, d, ep, x, w, chan, res, out, out_low, h->status, p.sat_pack, stream);
      if (cr < 0) return fail(cr, "%s", c1_last_error());
      if (cr == 0) { ++g_kernel_count[3]; return launch_check("conv1x1"); }
    }
    ++g_kernel_count[0];
    const int bn = (d->Cout % 128 == 0) ? 128 : 64;
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (d->w_layout == 1) p.w_tiled = w + (size_t)d->Cout * p.K;     // caller appended the re-tiled copy (hawq_retile_weights)
    else if (make_map_2d(&maps.b, w, (uint64_t)p.K, (uint64_t)d->Cout, (uint64_t)p.K, 64, (uint32_t)bn, CU_TENSOR_MAP_SWIZZLE_64B))
      return fail(HAWQ_ERR_CUDA, "hawq_conv2d: cuTensorMapEncodeTiled (weights) failed");
    p.tma_a = (!a4 && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0) ? 1 : 0;
    if (p.tma_a && make_map_2d(&maps.a, x, (uint64_t)d->Cin, (uint64_t)M, (uint64_t)d->Cin, 64, TC_BM, CU_TENSOR_MAP_SWIZZLE_64B))
      return fail(HAWQ_ERR_CUDA, "hawq_conv2d: cuTensorMapEncodeTiled (activations) failed");
    const bool wide = (d->Cout % 128 == 0);
    const long long tiles = ((M + TC_BM - 1) / TC_BM) * (d->Cout / (wide ? 128 : 64));
    const int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool patch_enabled = [] { const char* e = getenv("HAWQ_B200_PATCH"); return !(e && e[0] == '0'); }();
    if (patch_enabled && ep->mode == HAWQ_EPI_REQUANT && d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad == 1 && 128 + 2 * d->W + 2 <= 256) {
      const uint32_t rows = 128 + 2 * d->W + 2;
      const uint32_t rowb = a4 ? 32 : 64;
      if (make_map_2d(&maps.patch, x, (uint64_t)p.x_pix_bytes, (uint64_t)d->N * d->H * d->W, (uint64_t)p.x_pix_bytes, rowb, rows,
                      a4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B))
        return fail(HAWQ_ERR_CUDA, "hawq_conv2d: cuTensorMapEncodeTiled (input patch) failed");
      p.patch_rows = (int)rows;
    }
    if (ep->mode == HAWQ_EPI_RESIDUAL && ep->res_kind == 0 && ep->res_bits == 16 && ep->y_bits == 16) {
      const uint32_t cw = bn / 2;   // columns per epilogue warp
      const CUtensorMapSwizzle sw_y = cw * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
      int bad = make_map_2d(&maps.res, res, (uint64_t)d->Cout * 2, (uint64_t)M, (uint64_t)d->Cout * 2, cw * 2, 32, sw_y);
      bad |= make_map_2d(&maps.y, out, (uint64_t)d->Cout * 2, (uint64_t)M, (uint64_t)d->Cout * 2, cw * 2, 32, sw_y);
      if (ep->low_bits) {
        const uint32_t lb = cw * ep->low_bits / 8;   // low tile row bytes: 64 / 32 / 16
        const CUtensorMapSwizzle sw_l = lb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : lb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
        bad |= make_map_2d(&maps.low, out_low, (uint64_t)d->Cout * ep->low_bits / 8, (uint64_t)M, (uint64_t)d->Cout * ep->low_bits / 8, lb, 32, sw_l);
      }
      if (bad) return fail(HAWQ_ERR_CUDA, "hawq_conv2d: cuTensorMapEncodeTiled (residual epilogue) failed");
      p.tma_io = 1;
    }
    if (ep->mode == HAWQ_EPI_REQUANT) launch_tc<TC_EPI_REQ>(p, maps, wide, false, a4, grid, st);
    else if (ep->mode == HAWQ_EPI_RAW_I32) launch_tc<TC_EPI_RAW>(p, maps, wide, false, a4, grid, st);
    else {
      const int res_es = (ep->res_kind == 1 || ep->res_bits == 32) ? 4 : 2;
      if (res_es == 2 && ep->y_bits == 16) launch_tc<TC_EPI_RES22>(p, maps, wide, ratios_wide, a4, grid, st);
      else if (res_es == 4 && ep->y_bits == 32) launch_tc<TC_EPI_RES44>(p, maps, wide, ratios_wide, a4, grid, st);
      else if (res_es == 4 && ep->y_bits == 16) launch_tc<TC_EPI_RES42>(p, maps, wide, ratios_wide, a4, grid, st);
      else goto legacy;   // uint16 residual in, int32 out: not a combination the engine produces
    }
    return launch_check("conv_tc");
  }

legacy:
  const bool bn128 = (d->Cout % 128 == 0);
  const dim3 grid((unsigned)((M + CONV_BM - 1) / CONV_BM), (unsigned)(d->Cout / (bn128 ? 128 : 64)), 1);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->a_bits == 8) {
    if (bn128) launch_conv<128, false>(p, grid, s);
    else launch_conv<64, false>(p, grid, s);
  } else {
    if (bn128) launch_conv<128, true>(p, grid, s);
    else launch_conv<64, true>(p, grid, s);
  }
  return launch_check("conv_igemm");
}

// Resize-unit fusion: y = RHE(m1 * (conv1x1_s(x2, w2) + bias2)) + RHE(m * (conv1x1(x, w) + bias)), ReLU, uint16 stream +
// optional low-bit copy.  Both convolutions accumulate in TMEM inside one kernel (no int32 identity tensor in HBM).
int hawq_conv2d_dual(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x, const int8_t* w,
                     const hawq_chan* chan, const hawq_conv_desc* d2, const void* x2, const int8_t* w2, const hawq_chan* chan2,
                     void* out, void* out_low, void* stream) {
  if (!h || !d || !ep || !x || !w || !chan || !d2 || !x2 || !w2 || !chan2 || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_dual: null argument");
  static const bool tc_enabled = [] { const char* e = getenv("HAWQ_B200_TC"); return !(e && e[0] == '0'); }();
  static const bool dual_enabled = [] { const char* e = getenv("HAWQ_B200_DUAL"); return !(e && e[0] == '0'); }();
  if (!tc_enabled || !dual_enabled) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: disabled by environment");
  if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad != 0 || d2->kh != 1 || d2->kw != 1 || d2->pad != 0 || d2->stride < 1)
    return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: both convolutions must be 1x1 without padding (main stride 1)");
  if (d->N < 1 || d->H < 1 || d->W < 1 || d2->H < 1 || d2->W < 1) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_dual: bad geometry");
  if ((d->a_bits != 8 && d->a_bits != 4) || d2->a_bits != d->a_bits) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: a_bits must be equal and 4 or 8");
  if (d->Cin % 64 || d2->Cin % 64 || d->Cout % 64 || d2->Cout != d->Cout) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: channel counts must be multiples of 64 and Cout equal");
  if (d2->N != d->N || (d2->H - 1) / d2->stride + 1 != d->H || (d2->W - 1) / d2->stride + 1 != d->W)
    return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_dual: the two convolutions have different output grids");
  if (d->w_layout != 1 || d2->w_layout != 1) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: weights must carry the re-tiled copy (w_layout 1)");
  if (ep->mode != HAWQ_EPI_RESIDUAL || !ep->relu || ep->y_bits != 16) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: RESIDUAL + relu + uint16 stream only");
  if (ep->low_bits != 0 && ep->low_bits != 4 && ep->low_bits != 8) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_dual: low_bits must be 0/4/8");
  if (ep->low_bits) {
    if (!out_low) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_dual: low_bits set but out_low is null");
    int rc = check_me(ep->low_m, ep->low_e, "hawq_conv2d_dual low-bit copy");
    if (rc) return rc;
    if (!dyadic_is_fast(ep->low_m, ep->low_e) || !fold_ok(ep->low_m, ep->low_e)) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: low-bit ratio outside the fast range");
  }
  const bool ratios_one = (ep->flags & HAWQ_EP_RATIOS_LE_ONE) != 0;
  const bool ratios_wide = !ratios_one && (ep->flags & HAWQ_EP_RATIOS_LE_2P20) != 0;
  if (!ratios_one && !ratios_wide) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: needs a ratio-range promise (flags)");
  const long long M = (long long)d->N * d->H * d->W;
  if (M > 0x7fffff00ll || (long long)d2->N * d2->H * d2->W > 0x7fffff00ll) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_conv2d_dual: too many pixels");

  {   // int8 resize units: stationary weights, strided identity rows by TMA (conv_dual.cuh)
    const int dr = launch_conv_dual(h->sm_count, d, ep, x, w, chan, d2, x2, w2, chan2, out, out_low, h->status, sat_pack_enabled() ? 1 : 0, stream);
    if (dr < 0) return fail(dr, "%s", dual_last_error());
    if (dr == 0) { ++g_kernel_count[4]; return launch_check("conv_dual"); }
  }
  ++g_kernel_count[5];
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const uint8_t*)x; p.w = w; p.chan = chan; p.out = out; p.out_low = out_low; p.status = h->status;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.KH = 1; p.KW = 1; p.stride = 1; p.pad = 0;
  p.Ho = d->H; p.Wo = d->W; p.M = (int)M; p.K = d->Cin; p.cin_chunks = d->Cin / 64; p.x_pix_bytes = d->Cin * d->a_bits / 8;
  p.mode = ep->mode; p.relu = 1; p.res_kind = 1; p.y_bits = 16; p.low_bits = ep->low_bits; p.low_m = ep->low_m; p.low_e = ep->low_e;
  p.low_lo = ep->low_lo; p.low_hi = ep->low_hi;
  p.w_tiled = w + (size_t)d->Cout * d->Cin;
  p.dual = 1;
  p.sat_pack = sat_pack_enabled() ? 1 : 0;
  p.x2 = (const uint8_t*)x2; p.w2_tiled = w2 + (size_t)d2->Cout * d2->Cin; p.chan2 = chan2;
  p.H2 = d2->H; p.W2 = d2->W; p.stride2 = d2->stride; p.cin_chunks2 = d2->Cin / 64; p.x2_pix_bytes = d2->Cin * d2->a_bits / 8;
  p.tma_io = 1;

  const bool a4 = d->a_bits == 4;
  const bool bn128 = (d->Cout % 128 == 0);
  const uint32_t cw = (bn128 ? 128 : 64) / 2;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const CUtensorMapSwizzle sw_y = cw * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  int bad = make_map_2d(&maps.y, out, (uint64_t)d->Cout * 2, (uint64_t)M, (uint64_t)d->Cout * 2, cw * 2, 32, sw_y);
  if (ep->low_bits) {
    const uint32_t lb = cw * ep->low_bits / 8;
    const CUtensorMapSwizzle sw_l = lb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : lb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    bad |= make_map_2d(&maps.low, out_low, (uint64_t)d->Cout * ep->low_bits / 8, (uint64_t)M, (uint64_t)d->Cout * ep->low_bits / 8, lb, 32, sw_l);
  }
  if (bad) return fail(HAWQ_ERR_CUDA, "hawq_conv2d_dual: cuTensorMapEncodeTiled failed");
  const long long tiles = ((M + TC_BM - 1) / TC_BM) * (d->Cout / (bn128 ? 128 : 64));
  const int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
  launch_tc<TC_EPI_DUAL>(p, maps, bn128, ratios_wide, a4, grid, (cudaStream_t)stream);
  return launch_check("conv_tc_dual");
}

int hawq_conv2d_i8(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x,
                   const int8_t* w, const hawq_chan* chan, const void* res, const hawq_chan* res_chan,
                   const float* fscale, void* out, void* out_low, void* stream) {
  if (d && d->a_bits != 8) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_i8: a_bits must be 8");
  return hawq_conv2d(h, d, ep, x, w, chan, res, res_chan, fscale, out, out_low, stream);
}

int hawq_conv2d_i4(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep, const void* x,
                   const int8_t* w, const hawq_chan* chan, const void* res, const hawq_chan* res_chan,
                   const float* fscale, void* out, void* out_low, void* stream) {
  if (d && d->a_bits != 4) return fail(HAWQ_ERR_BAD_ARG, "hawq_conv2d_i4: a_bits must be 4");
  return hawq_conv2d(h, d, ep, x, w, chan, res, res_chan, fscale, out, out_low, stream);
}

int hawq_linear_i8(hawq_handle* h, int32_t N, int32_t K, int32_t Cout, int32_t Cout_pad, const int8_t* x,
                   const int8_t* w, const hawq_chan* chan, const float* fscale, float* out, void* stream) {
  if (Cout < 1 || Cout > Cout_pad) return fail(HAWQ_ERR_BAD_ARG, "hawq_linear_i8: bad Cout/Cout_pad");
  if (!h || !x || !w || !chan || !fscale || !out || N < 1 || K < 1) return fail(HAWQ_ERR_BAD_ARG, "hawq_linear_i8: null argument or empty shape");
  static const bool dp4a_enabled = [] { const char* e = getenv("HAWQ_B200_LINEAR_DP4A"); return !(e && e[0] == '0'); }();
  if (dp4a_enabled && K % LIN_SLAB == 0 && K <= LIN_MAX_K && Cout_pad % LIN_CH == 0 && N <= 65535 * LIN_ROWS) {
    const dim3 grid((unsigned)(Cout_pad / LIN_CH), (unsigned)((N + LIN_ROWS - 1) / LIN_ROWS), 1);
    linear_dp4a_kernel<<<grid, 256, linear_smem_bytes(K), (cudaStream_t)stream>>>(x, w, chan, fscale, out, N, K, Cout);
    return launch_check("linear_dp4a");
  }
  hawq_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.N = N; d.H = 1; d.W = 1; d.Cin = K; d.Cout = Cout_pad; d.kh = 1; d.kw = 1; d.stride = 1; d.pad = 0; d.a_bits = 8;
  hawq_epilogue_desc ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = HAWQ_EPI_DEQUANT_F32;
  ep.cout_store = Cout;
  return hawq_conv2d(h, &d, &ep, x, w, chan, nullptr, nullptr, fscale, out, nullptr, stream);
}

int hawq_stem_conv_i8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const int8_t* x, const int8_t* w,
                      const hawq_chan* chan, int32_t clamp_lo, int32_t clamp_hi, int16_t* out, void* stream) {
  if (!h || !x || !w || !chan || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_conv_i8: null argument");
  if (N < 1 || H < 7 || W < 7) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_conv_i8: bad geometry");
  if (clamp_lo < -32768 || clamp_hi > 32767 || clamp_lo > clamp_hi) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_conv_i8: clamp must fit int16");
  if (N > 65535) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_stem_conv_i8: N > 65535");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const dim3 grid((Wo + STEM_TW - 1) / STEM_TW, (Ho + STEM_TH - 1) / STEM_TH, N);
  // persistent: 4 CTAs per SM loop over the tiles and stage the weights once (round-2 A/B: 0.212 -> 0.202 ms at batch 128)
  const long long tiles = (long long)N * grid.x * grid.y;
  const int ctas = (int)(tiles < 4LL * h->sm_count ? tiles : 4LL * h->sm_count);
  stem_conv_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(x, (const uint32_t*)w, chan, N, H, W, Ho, Wo, clamp_lo, clamp_hi, out);
  return launch_check("stem_conv");
}

int hawq_stem_pool_i8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const int8_t* x, const int8_t* w256, const hawq_chan* chan,
                      int32_t clamp_lo, int32_t clamp_hi, int32_t y_bits, void* y, int32_t low_bits, uint32_t low_m, int32_t low_e,
                      int32_t low_lo, int32_t low_hi, void* out_low, void* stream) {
  if (!h || !x || !w256 || !chan || !y) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_pool_i8: null argument");
  if (N < 1 || H < 7 || W < 7) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_pool_i8: bad geometry");
  if (clamp_lo < -32768 || clamp_hi > 32767 || clamp_lo > clamp_hi) return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_pool_i8: clamp must fit int16");
  if ((y_bits != 16 && y_bits != 32) || (low_bits != 0 && low_bits != 4 && low_bits != 8) || (low_bits && !out_low))
    return fail(HAWQ_ERR_BAD_ARG, "hawq_stem_pool_i8: bad output description");
  if (low_bits) { int rc = check_me(low_m, low_e, "hawq_stem_pool_i8"); if (rc) return rc; }
  const int r = launch_stem_tc(h->sm_count, N, H, W, x, w256, chan, clamp_lo, clamp_hi, y_bits, y, low_bits, low_m, low_e, low_lo, low_hi, out_low,
                               h->status, stream);
  if (r < 0) return fail(r, "%s", stem_tc_last_error());
  if (r == 1) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_stem_pool_i8: shape / ratio outside the fused kernel (use hawq_stem_conv_i8 + hawq_maxpool_requant)");
  ++g_kernel_count[6];
  return launch_check("stem_tc");
}

int hawq_maxpool_requant(hawq_handle* h, int32_t N, int32_t H, int32_t W, int32_t C, const int16_t* x, int32_t y_bits,
                         void* y, int32_t low_bits, uint32_t low_m, int32_t low_e, int32_t low_lo, int32_t low_hi,
                         void* out_low, void* stream) {
  if (!h || !x) return fail(HAWQ_ERR_BAD_ARG, "hawq_maxpool_requant: null argument");
  if (C % 8 != 0) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_maxpool_requant: C %% 8 != 0");
  if ((y_bits != 0 && y_bits != 16 && y_bits != 32) || (y_bits && !y)) return fail(HAWQ_ERR_BAD_ARG, "hawq_maxpool_requant: bad y");
  if ((low_bits != 0 && low_bits != 4 && low_bits != 8) || (low_bits && !out_low)) return fail(HAWQ_ERR_BAD_ARG, "hawq_maxpool_requant: bad low");
  if (low_bits) { int rc = check_me(low_m, low_e, "hawq_maxpool_requant"); if (rc) return rc; }
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * Ho * Wo * (C / 8);
  maxpool_requant_kernel<<<grid_for(total, h->sm_count), 256, 0, (cudaStream_t)stream>>>(
      x, N, H, W, C, Ho, Wo, y_bits, y, low_bits, low_m, low_e, low_lo, low_hi, out_low);
  return launch_check("maxpool_requant");
}

int hawq_avgpool_requant(hawq_handle* h, int32_t N, int32_t HW, int32_t C, int32_t x_bits, const void* x, uint32_t m,
                         int32_t e, int32_t lo, int32_t hi, int8_t* out, void* stream) {
  if (!h || !x || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_avgpool_requant: null argument");
  if (x_bits != 16 && x_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_avgpool_requant: x_bits must be 16/32");
  if (lo < -128 || hi > 127) return fail(HAWQ_ERR_BAD_ARG, "hawq_avgpool_requant: clamp must fit int8");
  int rc = check_me(m, e, "hawq_avgpool_requant");
  if (rc) return rc;
  avgpool_requant_kernel<<<grid_for((long long)N * C, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, N, HW, C, x_bits, m, e, lo, hi, out);
  return launch_check("avgpool_requant");
}

int hawq_quantize_input_f32(hawq_handle* h, int32_t N, int32_t C, int32_t H, int32_t W, const float* x, float scale,
                            int32_t lo, int32_t hi, int8_t* out, void* stream) {
  if (!h || !x || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_f32: null argument");
  if (!(scale > 0.f)) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_f32: scale must be > 0");
  if (lo < -128 || hi > 127) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_f32: clamp must fit int8");
  const float inv = 1.0f / scale;  // fp32 division, as `1. / scale` in linear_quantize (quant_utils.py:97)
  quantize_input_kernel<<<grid_for((long long)N * H * W, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, N, C, H, W, inv, lo, hi, out);
  return launch_check("quantize_input");
}

int hawq_quantize_input_u8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const uint8_t* x, const float* mean3,
                           const float* std3, float scale, int32_t lo, int32_t hi, int8_t* out, void* stream) {
  if (!h || !x || !out || !mean3 || !std3) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_u8: null argument");
  if (N < 1 || H < 1 || W < 1) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_u8: empty shape");
  if (!(scale > 0.f)) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_u8: scale must be > 0");
  if (lo < -128 || hi > 127 || lo > hi) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_u8: clamp must fit int8");
  for (int c = 0; c < 3; ++c)
    if (!(std3[c] > 0.f)) return fail(HAWQ_ERR_BAD_ARG, "hawq_quantize_input_u8: std must be > 0");
  const float inv = 1.0f / scale;  // fp32 division, as `1. / scale` in linear_quantize (quant_utils.py:97)
  const long long n_bytes = (long long)N * H * W * 3;
  quantize_input_u8_kernel<<<grid_for(n_bytes / 4 + 1, h->sm_count), 256, 0, (cudaStream_t)stream>>>(
      x, n_bytes, 3, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], inv, lo, hi, out);
  return launch_check("quantize_input_u8");
}

int hawq_requant(hawq_handle* h, int64_t rows, int32_t C, int32_t x_bits, const void* x, const hawq_chan* chan,
                 int32_t chan_stride, int32_t relu, int32_t out_bits, int32_t lo, int32_t hi, void* out, void* stream) {
  if (!h || !x || !chan || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_requant: null argument");
  if (C % 8 != 0) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_requant: C %% 8 != 0");
  if (x_bits != 16 && x_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_requant: x_bits must be 16/32");
  if (out_bits != 4 && out_bits != 8 && out_bits != 16 && out_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_requant: out_bits must be 4/8/16/32");
  if (chan_stride != 0 && chan_stride != 1) return fail(HAWQ_ERR_BAD_ARG, "hawq_requant: chan_stride must be 0/1");
  requant_kernel<<<grid_for(rows * (C / 8), h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, rows, C, x_bits, chan, chan_stride, relu, out_bits, lo, hi, out);
  return launch_check("requant");
}

int hawq_add_requant(hawq_handle* h, int64_t rows, int32_t C, const int32_t* acc, const hawq_chan* chan,
                     const hawq_epilogue_desc* ep, const void* res, const hawq_chan* res_chan, void* y, void* out_low,
                     void* stream) {
  if (!h || !acc || !chan || !ep || !res) return fail(HAWQ_ERR_BAD_ARG, "hawq_add_requant: null argument");
  if (C % 8 != 0) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_add_requant: C %% 8 != 0");
  if (ep->res_kind == 1 && !res_chan) return fail(HAWQ_ERR_BAD_ARG, "hawq_add_requant: res_kind 1 needs res_chan");
  if (ep->res_kind == 0) { int rc = check_me(ep->res_m, ep->res_e, "hawq_add_requant"); if (rc) return rc; }
  if (ep->y_bits == 16 && !ep->relu) return fail(HAWQ_ERR_BAD_ARG, "hawq_add_requant: uint16 residual stream requires relu");
  if ((ep->y_bits && !y) || (ep->low_bits && !out_low)) return fail(HAWQ_ERR_BAD_ARG, "hawq_add_requant: missing output");
  AddRequantParams p;
  p.acc = acc; p.chan = chan; p.res = res; p.res_chan = res_chan; p.y = y; p.out_low = out_low; p.status = h->status;
  p.rows = rows; p.C = C; p.relu = ep->relu; p.res_kind = ep->res_kind; p.res_bits = ep->res_bits; p.res_m = ep->res_m;
  p.res_e = ep->res_e; p.y_bits = ep->y_bits; p.low_bits = ep->low_bits; p.low_m = ep->low_m; p.low_e = ep->low_e;
  p.low_lo = ep->low_lo; p.low_hi = ep->low_hi;
  add_requant_kernel<<<grid_for(rows * (C / 8), h->sm_count), 256, 0, (cudaStream_t)stream>>>(p);
  return launch_check("add_requant");
}

int hawq_dequant_f32(hawq_handle* h, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_bits, int32_t x_signed,
                     const void* x, float scale, float* out_nchw, void* stream) {
  if (!h || !x || !out_nchw) return fail(HAWQ_ERR_BAD_ARG, "hawq_dequant_f32: null argument");
  if (x_bits != 4 && x_bits != 8 && x_bits != 16 && x_bits != 32) return fail(HAWQ_ERR_BAD_ARG, "hawq_dequant_f32: bad x_bits");
  if (x_bits == 4 && C % 8 != 0) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_dequant_f32: packed input needs C %% 8 == 0");
  dequant_f32_kernel<<<grid_for((long long)N * H * W * C, h->sm_count), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, x_bits, x_signed, scale, out_nchw);
  return launch_check("dequant_f32");
}

int hawq_pack_i4(hawq_handle* h, int64_t n_values, const uint8_t* in, uint8_t* out, void* stream) {
  if (!h || !in || !out || n_values % 8 != 0) return fail(HAWQ_ERR_BAD_ARG, "hawq_pack_i4: bad argument");
  pack_i4_kernel<<<grid_for(n_values / 8, h->sm_count), 256, 0, (cudaStream_t)stream>>>(in, n_values / 8, out);
  return launch_check("pack_i4");
}

int hawq_unpack_i4(hawq_handle* h, int64_t n_values, const uint8_t* in, uint8_t* out, void* stream) {
  if (!h || !in || !out || n_values % 8 != 0) return fail(HAWQ_ERR_BAD_ARG, "hawq_unpack_i4: bad argument");
  unpack_i4_kernel<<<grid_for(n_values / 8, h->sm_count), 256, 0, (cudaStream_t)stream>>>(in, n_values / 8, out);
  return launch_check("unpack_i4");
}

// ------------------------------------------------------------------------------------------- host helpers
int hawq_dyadic(double ratio, uint32_t* m, int32_t* e) {
  if (!m || !e || !(ratio > 0.0) || !std::isfinite(ratio)) return fail(HAWQ_ERR_BAD_ARG, "hawq_dyadic: ratio must be positive and finite");
  int ex;
  const double mant = std::frexp(ratio, &ex);          // mant in [0.5, 1)
  const double scaled = mant * 2147483648.0;            // exact
  const double r = std::floor(scaled + 0.5);            // ROUND_HALF_UP for positive values, exact (see oracle/int_ref.py)
  const int ee = 31 - ex;
  if (ee < 1) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_dyadic: ratio %g too large (e = %d < 1)", ratio, ee);
  if (ee > 62) { *m = 0; *e = 1; return HAWQ_OK; }      // |v*m| < 2^62 <= 2^(e-1): always rounds to 0
  *m = (uint32_t)r;
  *e = ee;
  return HAWQ_OK;
}

int64_t hawq_rhe_requant_host(int32_t v, uint32_t m, int32_t e) { return (int64_t)rhe_requant(v, m, e); }

int hawq_permute_weights_for_i4(int8_t* host_w, int64_t rows_times_taps, int32_t Cin) {
  if (!host_w || Cin % 32 != 0) return fail(HAWQ_ERR_BAD_ARG, "hawq_permute_weights_for_i4: Cin %% 32 != 0");
  int8_t tmp[32];
  for (int64_t r = 0; r < rows_times_taps; ++r) {
    for (int blk = 0; blk < Cin / 32; ++blk) {
      int8_t* p = host_w + r * Cin + blk * 32;
      for (int t = 0; t < 4; ++t)
        for (int j = 0; j < 4; ++j) {
          tmp[4 * t + j] = p[8 * t + j];           // MMA k position 4t+j     <- channel 8t+j
          tmp[16 + 4 * t + j] = p[8 * t + 4 + j];  // MMA k position 16+4t+j  <- channel 8t+4+j
        }
      memcpy(p, tmp, 32);
    }
  }
  return HAWQ_OK;
}

int64_t hawq_workspace_bytes(const hawq_conv_desc*, const hawq_epilogue_desc*) { return 0; }

int hawq_retile_weights(hawq_handle* h, const int8_t* w_ohwi, int32_t Cout, int64_t K, int8_t* out, void* stream) {
  if (!h || !w_ohwi || !out) return fail(HAWQ_ERR_BAD_ARG, "hawq_retile_weights: null argument");
  if (Cout % 64 != 0 || K % 64 != 0) return fail(HAWQ_ERR_UNSUPPORTED, "hawq_retile_weights: Cout and K must be multiples of 64");
  const int bn = (Cout % 128 == 0) ? 128 : 64;
  const long long chunks = (long long)Cout * K / 16;
  retile_weights_kernel<<<grid_for(chunks, h->sm_count), 256, 0, (cudaStream_t)stream>>>(w_ohwi, Cout, (int)K, bn, out);
  return launch_check("retile_weights");
}

int32_t hawq_debug_halo_trace(int64_t* host_out, int32_t n) { return halo_read_trace(reinterpret_cast<long long*>(host_out), n); }
int32_t hawq_debug_c1_trace(int64_t* host_out, int32_t n) { return c1_read_trace(reinterpret_cast<long long*>(host_out), n); }

int64_t hawq_debug_kernel_count(int32_t family) { return (family >= 0 && family < 8) ? g_kernel_count[family] : -1; }

}  // extern "C"

/* hawq_b200.h — C ABI of libhawq_b200.so: the integer forward path of HAWQ-quantized ResNets on B200 (sm_100a).
 *
 * The reference (Zhen-Dong/HAWQ) has no FFI / operator-registration layer: its boundary for this path is the Python
 * nn.Module API of utils/quantization_utils/quant_modules.py.  Each entry point below therefore cites the reference
 * *module/function* whose frozen (eval) forward it replaces; hawq_b200/ (Python) mirrors the module API on top of
 * this ABI, INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless named host_*;
 *   - every call is asynchronous on the caller-supplied stream (cudaStream_t passed as void*), never allocates,
 *     never synchronises -> safe under CUDA-graph capture;
 *   - return value: 0 = HAWQ_OK, negative = hawq_status; hawq_last_error() gives a thread-local message;
 *   - activations are NHWC.  8-bit: one int8 per element.  4-bit: unsigned nibbles packed two per byte in the
 *     "hawq nibble order": inside every group of 8 consecutive channels, byte j (0..3) holds channel j in its low
 *     nibble and channel j+4 in its high nibble (so a 32-bit word expands to two int8x4 words with one AND and one
 *     SHIFT+AND).  hawq_pack_i4 / hawq_unpack_i4 convert from/to one-value-per-byte;
 *   - weights are int8, OHWI ([Cout][kh][kw][Cin], K-major) for 8- and 4-bit layers alike (Blackwell has no int4
 *     MMA: 4-bit weights are widened once at plan time; they are <1% of the traffic).  For layers whose INPUT is
 *     packed 4-bit the K order inside each 32-channel block must be permuted with hawq_permute_weights_for_i4
 *     (host helper) to match the on-chip nibble expansion;
 *   - the residual stream ("x16": the output of quant_act_int32, reference utils/models/q_resnet.py:120,254-258)
 *     is stored post-ReLU either as int32 (always exact) or as uint16 with saturation + sticky overflow flag
 *     (HAWQ_FLAG_RESIDUAL_OVERFLOW in the handle's status word; the host re-runs with int32 when it is set);
 *   - dyadic requantisation everywhere is q = RHE(v * m / 2^e), round-half-to-EVEN, i.e. the reference's
 *     torch.round(f64(v)*f64(m)/2^e) (utils/quantization_utils/quant_utils.py:406-408), with (m, e) from
 *     batch_frexp (quant_utils.py:188-213): 2^30 <= m <= 2^31, 1 <= e <= 62.
 */
#ifndef HAWQ_B200_H
#define HAWQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HAWQ_ABI_VERSION 1

typedef struct hawq_handle hawq_handle;

enum hawq_status {
  HAWQ_OK = 0,
  HAWQ_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, inconsistent descriptor */
  HAWQ_ERR_UNSUPPORTED = -2,  /* shape / bit-width combination this build has no kernel for */
  HAWQ_ERR_CUDA = -3          /* CUDA runtime error (message in hawq_last_error) */
};

/* bits of the device status word */
#define HAWQ_FLAG_RESIDUAL_OVERFLOW 1   /* a post-ReLU residual value exceeded 65535 while stored as uint16 */
#define HAWQ_FLAG_BAD_RATIO 2           /* a HAWQ_EP_RATIOS_* promise was broken: results invalid */
#define HAWQ_FLAG_REQUANT_OVERFLOW 4     /* a requantised value left int32 on the fast path (ratio > 1): re-run without HAWQ_EP_* flags */

/* Per-output-channel epilogue parameters (16 B, one vector load per channel).
 * bias = bias_integer (quant_modules.py:481-484), (m, e) = batch_frexp of the requant ratio of that channel. */
typedef struct {
  int32_t bias;
  uint32_t m;
  int32_t e;
  int32_t reserved;
} hawq_chan;

/* Convolution geometry.  Replaces the F.conv2d call of QuantBnConv2d.forward / QuantConv2d.forward
 * (quant_modules.py:493, :731-736).  Requirements: Cin % 64 == 0, Cout % 64 == 0 (pad on the host otherwise). */
typedef struct {
  int32_t N, H, W, Cin, Cout;
  int32_t kh, kw, stride, pad;
  int32_t a_bits;    /* 8: int8 NHWC input; 4: packed unsigned nibbles (hawq nibble order) */
  int32_t w_layout;  /* 0: w = OHWI only; 1: w = OHWI followed by the hawq_retile_weights copy (2 * Cout * K bytes) */
} hawq_conv_desc;

enum hawq_epilogue_mode {
  HAWQ_EPI_REQUANT = 0,   /* case 0, fixedpoint_fn (quant_utils.py:390-413): clamp(RHE((acc+bias)[relu] * m_c / 2^e_c)) */
  HAWQ_EPI_RESIDUAL = 1,  /* case 1 (quant_utils.py:416-456): RHE(res*m1/2^e1) + RHE((acc+bias)*m_c/2^e_c), no clamp, [relu] */
  HAWQ_EPI_RAW_I32 = 2,   /* acc + bias as int32 (identity-branch conv feeding case 1) */
  HAWQ_EPI_DEQUANT_F32 = 3 /* QuantLinear tail (quant_modules.py:129-130): float(acc+bias) * fscale[c] */
};

typedef struct {
  int32_t mode;           /* hawq_epilogue_mode */
  int32_t relu;           /* REQUANT: max(acc+bias,0) before requant; RESIDUAL: max(sum,0) after the add */
  /* REQUANT output */
  int32_t out_bits;       /* 4 (packed u4), 8 (int8), 16 (int16), 32 (int32) */
  int32_t clamp_lo, clamp_hi;
  /* RESIDUAL input operand */
  int32_t res_kind;       /* 0: residual-stream tensor, scalar (res_m, res_e); 1: int32 accumulator tensor, per-channel res_chan */
  int32_t res_bits;       /* res_kind 0: 16 (uint16) or 32 (int32) */
  uint32_t res_m;
  int32_t res_e;
  /* RESIDUAL outputs */
  int32_t y_bits;         /* 0: do not store the new residual stream; 16: uint16 (needs relu=1); 32: int32 */
  int32_t low_bits;       /* 0: none; 4 / 8: also store clamp(RHE(y * low_m / 2^low_e)) = the next quant_act's output */
  uint32_t low_m;
  int32_t low_e;
  int32_t low_lo, low_hi;
  /* DEQUANT_F32 */
  int32_t cout_store;     /* number of real output columns (<= Cout), row pitch of the fp32 output */
  int32_t flags;          /* HAWQ_EP_*: promises of the caller that unlock faster kernels */
} hawq_epilogue_desc;

/* Caller promise: every dyadic pair of this launch (chan[], res_chan[], res_m/e, low_m/e) has ratio m * 2^-e <= 1, i.e.
 * e >= 31 or m == 0 (true for every HAWQ ResNet layer).  Enables the tcgen05 kernel, which evaluates RHE(v * m / 2^e) with one
 * exact FP64 FMA.  The kernel re-checks the promise and raises HAWQ_FLAG_BAD_RATIO instead of computing wrong numbers. */
#define HAWQ_EP_RATIOS_LE_ONE 1
/* Weaker promise: every ratio <= 2^20 (e >= 11 or m == 0).  Same fast kernel plus an exact per-value check that the
 * requantised term fits int32; a violation raises HAWQ_FLAG_REQUANT_OVERFLOW (the generic kernels saturate instead). */
#define HAWQ_EP_RATIOS_LE_2P20 2

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
int hawq_abi_version(void);
const char* hawq_last_error(void);
int hawq_create(int device, hawq_handle** out);
int hawq_destroy(hawq_handle* h);
int hawq_sm_count(const hawq_handle* h);
/* status word (sticky flags set by kernels): reset is async on the stream, get synchronises the stream */
int hawq_reset_status(hawq_handle* h, void* stream);
int hawq_get_status(hawq_handle* h, void* stream, int32_t* host_flags);
/* async copy of the status word into a caller-owned device int32 (e.g. inside a CUDA graph) */
int hawq_copy_status(hawq_handle* h, int32_t* dst, void* stream);

/* ---- fused convolution (QuantBnConv2d / QuantConv2d + the QuantAct that consumes it) ------------------------- */
/* x: activations (int8 or packed u4, NHWC); w: int8 OHWI; chan[Cout];
 * res / res_chan: RESIDUAL operand (see res_kind); fscale[Cout]: DEQUANT_F32;
 * out: REQUANT result | RAW int32 | fp32 logits | new residual stream (RESIDUAL, y_bits != 0);
 * out_low: RESIDUAL low-bit copy (low_bits != 0). */
int hawq_conv2d(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep,
                const void* x, const int8_t* w, const hawq_chan* chan,
                const void* res, const hawq_chan* res_chan, const float* fscale,
                void* out, void* out_low, void* stream);
/* the two names SURVEY.md §8(b) proposes; thin checks over hawq_conv2d (a_bits must be 8 resp. 4) */
int hawq_conv2d_i8(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep,
                   const void* x, const int8_t* w, const hawq_chan* chan,
                   const void* res, const hawq_chan* res_chan, const float* fscale,
                   void* out, void* out_low, void* stream);
int hawq_conv2d_i4(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep,
                   const void* x, const int8_t* w, const hawq_chan* chan,
                   const void* res, const hawq_chan* res_chan, const float* fscale,
                   void* out, void* out_low, void* stream);

/* Resize residual units (Q_ResUnitBn with resize_identity, q_resnet.py:304-330): the identity-branch 1x1 convolution
 * (d2/x2/w2; stride d2->stride, pad 0) and the unit's last 1x1 convolution (d/x/w; stride 1) are accumulated side by side in
 * one kernel and combined by the case-1 fixed-point sum (quant_utils.py:430-456):
 *   y = ReLU( RHE((acc2 + chan2.bias) * chan2.m / 2^chan2.e) + RHE((acc + chan.bias) * chan.m / 2^chan.e) )
 * ep: mode RESIDUAL, relu 1, y_bits 16 (uint16 stream in out), optional low-bit copy in out_low, flags = ratio promise.
 * Both descriptors need w_layout 1 and equal a_bits / Cout / output grids.  Returns HAWQ_ERR_UNSUPPORTED for any other
 * combination (callers then use hawq_conv2d RAW_I32 followed by hawq_conv2d RESIDUAL res_kind 1: same results). */
int hawq_conv2d_dual(hawq_handle* h, const hawq_conv_desc* d, const hawq_epilogue_desc* ep,
                     const void* x, const int8_t* w, const hawq_chan* chan,
                     const hawq_conv_desc* d2, const void* x2, const int8_t* w2, const hawq_chan* chan2,
                     void* out, void* out_low, void* stream);

/* QuantLinear.forward (quant_modules.py:79-130): x int8 [N,K], w int8 [Cout_pad,K] (rows >= Cout zero),
 * chan[Cout_pad] (bias only), fscale[Cout_pad] = fc_scaling_factor[c] * act_scale (fp32 product) -> fp32 [N,Cout]. */
int hawq_linear_i8(hawq_handle* h, int32_t N, int32_t K, int32_t Cout, int32_t Cout_pad,
                   const int8_t* x, const int8_t* w, const hawq_chan* chan, const float* fscale,
                   float* out, void* stream);

/* ---- stem: 7x7 s2 p3 conv, Cin = 3 (Q_ResNet*.quant_init*_convbn, q_resnet.py:117) -------------------------- */
/* x int8 [N,H,W,3]; w int8 [64][7][8][4] (kw 7 and channel 3 zero); chan[64] carries bias and the 16-bit requant of
 * quant_act_int32 (q_resnet.py:120).  Output int16 [N,Ho,Wo,64] = max(0, clamp(RHE((acc+bias)*m/2^e), lo, hi)):
 * requant and ReLU commute with the max-pool that follows (both monotone). */
int hawq_stem_conv_i8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const int8_t* x, const int8_t* w,
                      const hawq_chan* chan, int32_t clamp_lo, int32_t clamp_hi, int16_t* out, void* stream);

/* Fused stem (tcgen05): quant_init_convbn (7x7 stride 2 pad 3, Cin = 3 -> 64) + nn.MaxPool2d(3, 2, 1) + quant_act_int32 (16-bit dyadic
 * requant, clamp) + ReLU, and optionally the first unit's low-bit quant_act (q_resnet.py:117-122, :234) in one kernel; the int16
 * convolution output never reaches HBM.  w256 = int8 [64][8][8][4] (kernel rows padded 7 -> 8, taps 7 -> 8, channels 3 -> 4, zeros
 * in the padding).  y = pooled residual stream [N][Hp][Wp][64] as uint16 (y_bits 16) or int32 (32).  Preconditions: every ratio
 * <= 1, W % 4 == 0, W <= 256; otherwise HAWQ_ERR_UNSUPPORTED (use hawq_stem_conv_i8 + hawq_maxpool_requant: same integers). */
int hawq_stem_pool_i8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const int8_t* x, const int8_t* w256, const hawq_chan* chan,
                      int32_t clamp_lo, int32_t clamp_hi, int32_t y_bits, void* y, int32_t low_bits, uint32_t low_m, int32_t low_e,
                      int32_t low_lo, int32_t low_hi, void* out_low, void* stream);
/* nn.MaxPool2d(3,2,1) (q_resnet.py:119) on the int16 stem output + the first unit's quant_act (case 0, scalar m,e).
 * y: residual stream (y_bits 16 -> uint16, 32 -> int32); out_low: int8 / packed u4 (low_bits 8 / 4, 0 = none). */
int hawq_maxpool_requant(hawq_handle* h, int32_t N, int32_t H, int32_t W, int32_t C, const int16_t* x,
                         int32_t y_bits, void* y, int32_t low_bits, uint32_t low_m, int32_t low_e,
                         int32_t low_lo, int32_t low_hi, void* out_low, void* stream);

/* QuantAveragePool2d (quant_modules.py:585-602) + quant_act_output (q_resnet.py:131): x residual stream
 * [N,HW,C] (x_bits 16/32) -> int8 [N,C] = clamp(RHE(trunc_avg(x) * m / 2^e)). */
int hawq_avgpool_requant(hawq_handle* h, int32_t N, int32_t HW, int32_t C, int32_t x_bits, const void* x,
                         uint32_t m, int32_t e, int32_t lo, int32_t hi, int8_t* out, void* stream);

/* ---- stand-alone (unfused) pieces of the module API --------------------------------------------------------- */
/* QuantAct input branch (quant_modules.py:271-274): q = clamp(round((1/scale) * x)), fp32 RNE.
 * x fp32 NCHW [N,C,H,W] -> int8 NHWC [N,H,W,C]. */
int hawq_quantize_input_f32(hawq_handle* h, int32_t N, int32_t C, int32_t H, int32_t W, const float* x,
                            float scale, int32_t lo, int32_t hi, int8_t* out, void* stream);
/* uint8 image entry (tvm_benchmark/test_resnet_accuracy_imagenet.py:62-75 quantize_image after transforms.ToTensor + Normalize
 * :82-93): x uint8 NHWC [N,H,W,3] -> int8 NHWC, q = clamp(round((1/scale) * ((x / 255 - mean[c]) / std[c]))), every step one fp32
 * operation as in the torch pipeline.  mean3 / std3 are HOST pointers to three floats (copied at launch). */
int hawq_quantize_input_u8(hawq_handle* h, int32_t N, int32_t H, int32_t W, const uint8_t* x, const float* mean3,
                           const float* std3, float scale, int32_t lo, int32_t hi, int8_t* out, void* stream);
/* fixedpoint_fn case 0 stand-alone (QuantAct after a conv or at unit entry): x [rows,C] (x_bits 16 = uint16 residual,
 * 32 = int32), per-channel chan (bias is added; pass 0) or scalar when chan_stride == 0 (chan[0] used for all). */
int hawq_requant(hawq_handle* h, int64_t rows, int32_t C, int32_t x_bits, const void* x, const hawq_chan* chan,
                 int32_t chan_stride, int32_t relu, int32_t out_bits, int32_t lo, int32_t hi, void* out, void* stream);
/* fixedpoint_fn case 1 stand-alone: y = [relu](RHE(res*m1/2^e1) + RHE((acc+bias)*m/2^e)); same operands as the fused form. */
int hawq_add_requant(hawq_handle* h, int64_t rows, int32_t C, const int32_t* acc, const hawq_chan* chan,
                     const hawq_epilogue_desc* ep, const void* res, const hawq_chan* res_chan,
                     void* y, void* out_low, void* stream);
/* integer tensor -> fp32 NCHW "fake-quant" value q * scale (graph edges of the module API). x_bits 4 (packed), 8, 16 (uint16), 32. */
int hawq_dequant_f32(hawq_handle* h, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_bits, int32_t x_signed,
                     const void* x, float scale, float* out_nchw, void* stream);
/* one value per byte (0..15) <-> packed nibbles in hawq nibble order; n_values % 8 == 0 */
int hawq_pack_i4(hawq_handle* h, int64_t n_values, const uint8_t* in, uint8_t* out, void* stream);
int hawq_unpack_i4(hawq_handle* h, int64_t n_values, const uint8_t* in, uint8_t* out, void* stream);

/* ---- host helpers (no GPU needed) --------------------------------------------------------------------------- */
/* batch_frexp (quant_utils.py:188-213) of one positive ratio: m = round_half_up(mant * 2^31), e = 31 - exp.
 * Returns HAWQ_ERR_UNSUPPORTED when e < 1 (ratio >= 2^30); for e > 62 the result is always 0: (m, e) := (0, 1). */
int hawq_dyadic(double ratio, uint32_t* m, int32_t* e);
/* exact host evaluation of RHE(v * m / 2^e) — the same routine the kernels inline */
int64_t hawq_rhe_requant_host(int32_t v, uint32_t m, int32_t e);
/* K permutation inside each 32-channel block for layers whose input is packed 4-bit (in place, int8 OHWI, host memory) */
int hawq_permute_weights_for_i4(int8_t* host_w, int64_t rows_times_taps, int32_t Cin);
/* Re-tile int8 OHWI weights [Cout][K] for the tcgen05 convolution: block (n_tile, k_tile) = BN rows x 64 bytes (BN = 128 when
 * Cout % 128 == 0, else 64), stored contiguously with the shared-memory swizzle pre-applied, so the kernel fetches a k-tile of
 * weights with one linear bulk copy.  `out` (Cout * K bytes) is normally w_ohwi + Cout * K, i.e. the copy is appended to the OHWI
 * tensor and announced with hawq_conv_desc.w_layout = 1.  Device pointers, asynchronous on the stream. */
int hawq_retile_weights(hawq_handle* h, const int8_t* w_ohwi, int32_t Cout, int64_t K, int8_t* out, void* stream);
/* debug: number of hawq_conv2d launches so far that went to kernel family 0 = conv_tc (generic tcgen05 implicit GEMM),
 * 1 = conv_halo (3x3 stride-1, A operand read in place), 2 = conv_halo launches that needed the 2-D weight-map fallback,
 * 3 = conv1x1 (1x1 stride-1, stationary weights), 4 = conv_dual (resize-unit tail, stationary weights), 5 = resize-unit tails
 * on conv_tc, 6 = fused tcgen05 stem;
 * -1 for an unknown family.  Lets tests assert which kernel ran. */
int64_t hawq_debug_kernel_count(int32_t family);
/* debug: with HAWQ_B200_HALO_TRACE=1 in the environment every conv_halo launch records clock64 stamps of CTA 0
 * ([3 roles: producer, MMA issuer, epilogue][64 steps][4 events]); this copies the last launch's buffer to the host (synchronises
 * the device) and returns the number of int64 values written.  Not for production use. */
int32_t hawq_debug_halo_trace(int64_t* host_out, int32_t n);
/* same for the last conv1x1 launch: [4 roles: producer / converter, MMA issuer, epilogue, residual loader][48 tiles][4 events] */
int32_t hawq_debug_c1_trace(int64_t* host_out, int32_t n);
/* workspace query kept for ABI completeness: this build needs no scratch beyond caller tensors */
int64_t hawq_workspace_bytes(const hawq_conv_desc* d, const hawq_epilogue_desc* ep);

#ifdef __cplusplus
}
#endif
#endif /* HAWQ_B200_H */

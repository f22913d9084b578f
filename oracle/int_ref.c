/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer semantics of HAWQ's quantized forward path.
 *
 * A second, independent oracle next to oracle/int_ref.py (numpy): same definitions, different language and arithmetic (pure
 * 64-bit integers, no floating point except where the reference itself uses it).  Only tests/ may load it; the product never
 * does.  Built by __graft_entry__.build() / tests into oracle/_build/libintref.so (gcc -O2 -shared -fPIC).
 *
 * Reference (Zhen-Dong/HAWQ) definitions followed, file:line in the reference checkout:
 *   batch_frexp                      utils/quantization_utils/quant_utils.py:188-213
 *   fixedpoint_fn case 0 / case 1    utils/quantization_utils/quant_utils.py:390-413 / 416-456   (torch.round = half-to-even)
 *   clamp ranges                     utils/quantization_utils/quant_utils.py:365-368
 *   QuantAct input quantisation      utils/quantization_utils/quant_modules.py:271-274 + quant_utils.py:73-97
 *   QuantBnConv2d / QuantConv2d      utils/quantization_utils/quant_modules.py:493, 731-736 (integer convolution + integer bias)
 *   QuantAveragePool2d               utils/quantization_utils/quant_modules.py:585-602 + quant_utils.py:324-341
 *   nn.MaxPool2d(3, 2, 1)            utils/models/q_resnet.py:93,119
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* batch_frexp for one positive double: m = ROUND_HALF_UP(mantissa * 2^31), e = 31 - exponent; m may be 2^31. */
void hawq_ref_dyadic(double r, int64_t* m, int32_t* e) {
  int ex;
  const double mant = frexp(r, &ex);          /* r = mant * 2^ex, 0.5 <= mant < 1 */
  const double x = ldexp(mant, 31);           /* exact */
  *m = (int64_t)floor(x + 0.5);               /* x > 0: half-up; exact because x < 2^31 leaves >= 22 fractional bits */
  *e = 31 - ex;
}

/* round-half-to-even of p / 2^e, 1 <= e <= 62, p any int64 */
int64_t hawq_ref_rhe_shift(int64_t p, int32_t e) {
  const int64_t q = p >> e;                   /* floor (arithmetic shift) */
  const int64_t rem = p - (q << e);           /* 0 <= rem < 2^e */
  const int64_t half = (int64_t)1 << (e - 1);
  if (rem > half || (rem == half && (q & 1))) return q + 1;
  return q;
}

static int64_t sat32(int64_t v) { return v > 2147483647LL ? 2147483647LL : (v < -2147483648LL ? -2147483648LL : v); }

/* RHE(v * m / 2^e) for |v| < 2^31, m <= 2^31 */
int64_t hawq_ref_requant1(int64_t v, int64_t m, int32_t e) { return hawq_ref_rhe_shift(v * m, e); }

/* fixedpoint_fn case 0: out[i][c] = clamp(RHE(([relu](acc + bias[c])) * m[c] / 2^e[c]), lo, hi); rows x C, per-channel m/e
 * (per_channel = 0: m[0], e[0] for every channel).  bias may be NULL. */
void hawq_ref_requant_case0(const int64_t* acc, int64_t rows, int32_t C, const int64_t* bias, const int64_t* m, const int32_t* e,
                            int32_t per_channel, int32_t relu, int64_t lo, int64_t hi, int64_t* out) {
  for (int64_t r = 0; r < rows; ++r)
    for (int32_t c = 0; c < C; ++c) {
      int64_t v = acc[r * C + c] + (bias ? bias[c] : 0);
      if (relu && v < 0) v = 0;
      int64_t q = hawq_ref_requant1(v, m[per_channel ? c : 0], e[per_channel ? c : 0]);
      out[r * C + c] = q < lo ? lo : (q > hi ? hi : q);
    }
}

/* fixedpoint_fn case 1 (unclamped): out = [relu](RHE(id * m1 / 2^e1) + RHE(acc * m2 / 2^e2)); m1/e1 scalar or per channel */
void hawq_ref_requant_case1(const int64_t* acc, const int64_t* id, int64_t rows, int32_t C, const int64_t* m2, const int32_t* e2,
                            const int64_t* m1, const int32_t* e1, int32_t id_per_channel, int32_t relu, int64_t* out) {
  for (int64_t r = 0; r < rows; ++r)
    for (int32_t c = 0; c < C; ++c) {
      const int64_t a = hawq_ref_requant1(acc[r * C + c], m2[c], e2[c]);
      const int64_t b = hawq_ref_requant1(id[r * C + c], m1[id_per_channel ? c : 0], e1[id_per_channel ? c : 0]);
      int64_t y = a + b;
      if (relu && y < 0) y = 0;
      out[r * C + c] = y;
    }
}

/* integer convolution, NHWC activations x [N,H,W,Cin], OHWI weights w [Cout,kh,kw,Cin], zero padding, + integer bias (may be NULL)
 * -> out [N,Ho,Wo,Cout] (int64 accumulators; the reference keeps them in fp32 tensors holding integers) */
void hawq_ref_conv2d_nhwc(const int32_t* x, const int32_t* w, const int64_t* bias, int32_t N, int32_t H, int32_t W, int32_t Cin,
                          int32_t Cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int64_t* out) {
  const int32_t Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int32_t n = 0; n < N; ++n)
    for (int32_t ho = 0; ho < Ho; ++ho)
      for (int32_t wo = 0; wo < Wo; ++wo)
        for (int32_t co = 0; co < Cout; ++co) {
          int64_t acc = bias ? bias[co] : 0;
          for (int32_t i = 0; i < kh; ++i) {
            const int32_t hi = ho * stride - pad + i;
            if (hi < 0 || hi >= H) continue;
            for (int32_t j = 0; j < kw; ++j) {
              const int32_t wi = wo * stride - pad + j;
              if (wi < 0 || wi >= W) continue;
              const int32_t* xp = x + (((int64_t)n * H + hi) * W + wi) * Cin;
              const int32_t* wp = w + (((int64_t)co * kh + i) * kw + j) * Cin;
              for (int32_t c = 0; c < Cin; ++c) acc += (int64_t)xp[c] * wp[c];
            }
          }
          out[(((int64_t)n * Ho + ho) * Wo + wo) * Cout + co] = acc;
        }
}

/* nn.MaxPool2d(3, 2, 1) on NHWC integers (padding never wins: -inf) */
void hawq_ref_maxpool_3x3_s2_p1(const int64_t* x, int32_t N, int32_t H, int32_t W, int32_t C, int64_t* out) {
  const int32_t Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  for (int32_t n = 0; n < N; ++n)
    for (int32_t ho = 0; ho < Ho; ++ho)
      for (int32_t wo = 0; wo < Wo; ++wo)
        for (int32_t c = 0; c < C; ++c) {
          int64_t best = INT64_MIN;
          for (int32_t i = 0; i < 3; ++i)
            for (int32_t j = 0; j < 3; ++j) {
              const int32_t hi = 2 * ho - 1 + i, wi = 2 * wo - 1 + j;
              if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
              const int64_t v = x[(((int64_t)n * H + hi) * W + wi) * C + c];
              if (v > best) best = v;
            }
          out[(((int64_t)n * Ho + ho) * Wo + wo) * C + c] = best;
        }
}

/* QuantAveragePool2d over the whole k x k map: trunc(sum / k^2 + 0.01) evaluated as the reference does (fp32 mean, +0.01, trunc)
 * restated in integers: floor for sums >= 0; toward zero for negative sums, except exact multiples, which lose one (-q -> -q + 1) */
void hawq_ref_avgpool_trunc(const int64_t* x, int32_t N, int32_t HW, int32_t C, int64_t* out) {
  for (int32_t n = 0; n < N; ++n)
    for (int32_t c = 0; c < C; ++c) {
      int64_t s = 0;
      for (int32_t k = 0; k < HW; ++k) s += x[((int64_t)n * HW + k) * C + c];
      int64_t r;
      if (s >= 0) r = s / HW;
      else {
        const int64_t a = -s, q = a / HW;
        r = (a % HW == 0) ? -q + 1 : -q;
      }
      out[(int64_t)n * C + c] = r;
    }
}

/* QuantAct input branch: q = clamp(rint((1 / scale) * x)) in fp32 (round-half-even), NCHW fp32 -> NHWC integers */
void hawq_ref_quantize_input(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float scale, int64_t lo, int64_t hi, int64_t* out) {
  const float inv = 1.0f / scale;
  for (int32_t n = 0; n < N; ++n)
    for (int32_t c = 0; c < C; ++c)
      for (int32_t h = 0; h < H; ++h)
        for (int32_t w = 0; w < W; ++w) {
          volatile float prod = inv * x[(((int64_t)n * C + c) * H + h) * W + w];   /* volatile: one fp32 rounding, no contraction */
          int64_t q = (int64_t)rintf(prod);
          out[(((int64_t)n * H + h) * W + w) * C + c] = q < lo ? lo : (q > hi ? hi : q);
        }
}

int32_t hawq_ref_sat32_selftest(void) { return sat32(1LL << 40) == 2147483647LL && sat32(-(1LL << 40)) == -2147483648LL; }

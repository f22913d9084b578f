// Implicit-GEMM integer convolution with fused HAWQ epilogues (stage-A kernel: IMMA m16n8k32 + cp.async pipeline).
//
//   M = N*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin.   A[m,k] gathered on the fly from the NHWC activation
//   tensor (zero-filled outside the image), B = int8 OHWI weights.  CTA tile 128 x BN x 64 channels, 8 warps
//   (4 along M x 2 along N), 4-stage cp.async ring, XOR-swizzled shared memory read with ldmatrix.
//
//   A4 = true: activations are packed unsigned nibbles (hawq nibble order).  They stay packed in HBM and in shared
//   memory (half the bytes); each ldmatrix word (8 nibbles) is expanded in registers with AND / SHIFT+AND into the
//   two int8x4 words of the MMA A fragment.  The weight rows of such layers are K-permuted on the host so that the
//   expansion needs no shuffles (hawq_permute_weights_for_i4).
//
//   Epilogues (hawq_epilogue_mode): REQUANT (case 0 of fixedpoint_fn), RESIDUAL (case 1: dual dyadic requant + add,
//   optional ReLU, writes the new residual stream and/or the next unit's low-bit activation), RAW_I32, DEQUANT_F32.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hawq {

struct ConvParams {
  const uint8_t* x;
  const int8_t* w;
  const hawq_chan* chan;
  const void* res;
  const hawq_chan* res_chan;
  const float* fscale;
  void* out;
  void* out_low;
  int32_t* status;
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, M, K;
  int cin_chunks;   // Cin / 64
  int x_pix_bytes;  // bytes per input pixel (Cin * a_bits / 8)
  int mode, relu, out_bits, lo, hi;
  int res_kind, res_bits;
  uint32_t res_m;
  int res_e;
  int y_bits, low_bits;
  uint32_t low_m;
  int low_e, low_lo, low_hi;
  int cout_store;
  int slow_scalar;   // host-checked: a scalar dyadic pair (res / low) has ratio > 1 -> generic 64-bit requant
  int tma_a;         // tcgen05 kernel: activations are fetched by TMA (1x1 stride-1 int8 layers)
  int tma_io;        // tcgen05 kernel: uint16 residual tile in / outputs out through TMA
  const int8_t* w_tiled;  // tcgen05 kernel: weights re-tiled into contiguous pre-swizzled [n_tile][k_tile][BN][64] blocks
  int patch_rows;    // tcgen05 kernel, 3x3 stride-1 pad-1 layers: rows of the shared-memory input patch (128 + 2W + 2), 0 = gather mode
  int sat_pack;           // tcgen05 RESIDUAL epilogues: 8-bit low copy packed with cvt.pack.sat when its clamp is [<= 0, 127]
  // tcgen05 dual mode (resize units): the identity-branch 1x1 convolution is computed in the same kernel into a second
  // TMEM accumulator instead of round-tripping an int32 tensor through HBM
  int dual;               // 0 / 1
  const uint8_t* x2;      // identity conv input (same a_bits as x)
  const int8_t* w2_tiled; // its re-tiled weights
  const hawq_chan* chan2; // its bias and the case-1 identity ratio (m1, e1) per channel
  int H2, W2, stride2, cin_chunks2, x2_pix_bytes;
};

constexpr int CONV_BM = 128;
constexpr int CONV_STAGES = 4;
constexpr int CONV_THREADS = 256;

template <int BN, bool A4>
struct ConvSmem {
  static constexpr int A_ROW = A4 ? 32 : 64;  // bytes of one A row per k-tile (64 channels)
  static constexpr int A_STAGE = CONV_BM * A_ROW;
  static constexpr int B_STAGE = BN * 64;
  static constexpr int PIPE = CONV_STAGES * (A_STAGE + B_STAGE);
  static constexpr int OUT_PITCH = BN + 16;
  static constexpr int OUT_STAGE = CONV_BM * OUT_PITCH;
  static constexpr int RES_TILE = CONV_BM * (BN * 4 + 32);                 // residual tile, worst case int32 + padding
  static constexpr int MAIN = PIPE > RES_TILE ? PIPE : RES_TILE;          // pipeline ring, later the residual / y tile
  static constexpr int OUT_OFF = MAIN;                                     // low-bit output staging tile
  static constexpr int CHAN_OFF = OUT_OFF + OUT_STAGE;                     // hawq_chan[BN]
  static constexpr int M_OFF = CHAN_OFF + BN * (int)sizeof(hawq_chan);     // double[BN]: m * 2^-e of chan
  static constexpr int M1_OFF = M_OFF + BN * 8;                           // double[BN]: m * 2^-e of res_chan
  static constexpr int RC_OFF = M1_OFF + BN * 8;                          // hawq_chan[BN]: res_chan
  static constexpr int CB_OFF = RC_OFF + BN * (int)sizeof(hawq_chan);     // double[BN]: 2^52 + 2^31 - bias
  static constexpr int TOTAL = CB_OFF + BN * 8;
};

// swizzled byte offset of 16-byte chunk `ch` of row `row` (rows of 64 B: 4 chunks; rows of 32 B: 2 chunks)
template <int ROW_BYTES>
__device__ __forceinline__ int swz(int row, int ch) {
  if constexpr (ROW_BYTES == 64) return row * 64 + ((ch ^ ((row >> 1) & 3)) << 4);
  else return row * 32 + ((ch ^ ((row >> 2) & 1)) << 4);
}

// EPI selects the compile-time specialised fast epilogue (used when every dyadic ratio of the launch is <= 1, which the
// kernel verifies): 0 = none (generic run-time epilogue only), 1 = REQUANT to 4/8 bits, 2 = RESIDUAL.
constexpr int EPI_GENERIC = 0, EPI_FAST_LOW = 1, EPI_FAST_RES = 2;

template <int BN, bool A4, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 2) conv_igemm_kernel(const ConvParams p) {
  using S = ConvSmem<BN, A4>;
  constexpr int BM = CONV_BM, STAGES = CONV_STAGES;
  constexpr int A_ROW = S::A_ROW;
  constexpr int A_CH = A_ROW / 16;                    // 16-byte chunks per A row: 4 or 2
  constexpr int A_ROWS_PER_PASS = CONV_THREADS / A_CH;  // 64 or 128
  constexpr int A_PASSES = BM / A_ROWS_PER_PASS;        // 2 or 1
  constexpr int B_PASSES = BN / 64;
  constexpr int WNT = BN / 2;   // warp tile width
  constexpr int NT = WNT / 8;   // n8 tiles per warp: 8 or 4

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * S::A_STAGE;
  hawq_chan* sChan = reinterpret_cast<hawq_chan*>(smem + S::CHAN_OFF);
  double* sM = reinterpret_cast<double*>(smem + S::M_OFF);
  double* sM1 = reinterpret_cast<double*>(smem + S::M1_OFF);
  hawq_chan* sResChan = reinterpret_cast<hawq_chan*>(smem + S::RC_OFF);
  double* sCb = reinterpret_cast<double*>(smem + S::CB_OFF);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  int slow = p.slow_scalar;
  if (tid < BN) {
    const hawq_chan c = p.chan[n0 + tid];
    sChan[tid] = c;
    sM[tid] = dyadic_to_double(c.m, c.e);
    sCb[tid] = 4503601774854144.0 - (double)c.bias;   // exact: folds the bias add into the int -> double conversion
    slow |= !dyadic_is_fast(c.m, c.e);
    if (p.mode == HAWQ_EPI_RESIDUAL && p.res_kind == 1) {
      const hawq_chan rc = p.res_chan[n0 + tid];
      sResChan[tid] = rc;
      sM1[tid] = dyadic_to_double(rc.m, rc.e);
      slow |= !dyadic_is_fast(rc.m, rc.e);
    }
  }
  const bool use_slow = __syncthreads_or(slow) != 0;   // CTA-uniform: any ratio > 1 -> generic exact integer requant

  // ---- per-thread gather coordinates for the A rows this thread copies ----
  const int a_ch = tid % A_CH;
  int a_hi0[A_PASSES], a_wi0[A_PASSES], a_pix[A_PASSES];
  bool a_ok[A_PASSES];
#pragma unroll
  for (int i = 0; i < A_PASSES; ++i) {
    const int row = tid / A_CH + i * A_ROWS_PER_PASS;
    const int m = m0 + row;
    a_ok[i] = m < p.M;
    const int mm = a_ok[i] ? m : 0;
    const int n = mm / (p.Ho * p.Wo);
    const int r = mm - n * (p.Ho * p.Wo);
    const int ho = r / p.Wo, wo = r - ho * p.Wo;
    a_hi0[i] = ho * p.stride - p.pad;
    a_wi0[i] = wo * p.stride - p.pad;
    a_pix[i] = n * p.H * p.W;
  }
  const int b_ch = tid & 3, b_row = tid >> 2;
  const int KT = p.KH * p.KW * p.cin_chunks;
  int ld_c = 0, ld_kw = 0, ld_kh = 0, ld_kt = 0;

  auto load_tile = [&](int stage) {
    const uint32_t a_base = smem_u32(sA + stage * S::A_STAGE);
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i) {
      const int row = tid / A_CH + i * A_ROWS_PER_PASS;
      const int hi = a_hi0[i] + ld_kh, wi = a_wi0[i] + ld_kw;
      const bool v = a_ok[i] && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W;
      const uint8_t* src = p.x;
      if (v) src = p.x + (size_t)(a_pix[i] + hi * p.W + wi) * p.x_pix_bytes + ld_c * A_ROW + a_ch * 16;
      cp_async_16(a_base + swz<A_ROW>(row, a_ch), src, v ? 16 : 0);
    }
    const uint32_t b_base = smem_u32(sB + stage * S::B_STAGE);
#pragma unroll
    for (int i = 0; i < B_PASSES; ++i) {
      const int row = b_row + i * 64;
      const int8_t* src = p.w + (size_t)(n0 + row) * p.K + ld_kt * 64 + b_ch * 16;
      cp_async_16(b_base + swz<64>(row, b_ch), src, 16);
    }
    ++ld_kt;
    if (++ld_c == p.cin_chunks) {
      ld_c = 0;
      if (++ld_kw == p.KW) { ld_kw = 0; ++ld_kh; }
    }
  };

  int32_t acc[2][NT][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[mi][ni][k] = 0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < KT) load_tile((kt + STAGES - 1) % STAGES);
    cp_async_commit();

    const int stage = kt % STAGES;
    const uint32_t a_base = smem_u32(sA + stage * S::A_STAGE);
    const uint32_t b_base = smem_u32(sB + stage * S::B_STAGE);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int row = wm * 32 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        if constexpr (!A4) {
          ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], a_base + swz<64>(row, ks * 2 + (lane >> 4)));
        } else {
          uint32_t r0, r1;
          ldmatrix_x2(r0, r1, a_base + swz<32>(row, ks));
          af[mi][0] = r0 & 0x0F0F0F0Fu;
          af[mi][1] = r1 & 0x0F0F0F0Fu;
          af[mi][2] = (r0 >> 4) & 0x0F0F0F0Fu;
          af[mi][3] = (r1 >> 4) & 0x0F0F0F0Fu;
        }
      }
      uint32_t bf[NT][2];
#pragma unroll
      for (int nj = 0; nj < NT / 2; ++nj) {
        const int row = wn * WNT + nj * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(bf[2 * nj][0], bf[2 * nj][1], bf[2 * nj + 1][0], bf[2 * nj + 1][1],
                    b_base + swz<64>(row, ks * 2 + ((lane >> 3) & 1)));
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) mma_16832<A4>(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // pipeline buffers are free: reused for the residual tile (RESIDUAL epilogue)

  // RESIDUAL: bulk-load this tile of the residual operand (coalesced 16 B cp.async, zero-fill past M) instead of
  // issuing dependent scalar loads inside the epilogue; padded pitch keeps the fragment-pattern reads conflict-free.
  uint8_t* sRes = smem;
  const int res_es = (p.mode == HAWQ_EPI_RESIDUAL) ? ((p.res_kind == 1 || p.res_bits == 32) ? 4 : 2) : 0;
  const int res_pitch = BN * res_es + 8 * res_es;
  if (res_es) {
    const int cpr = BN * res_es / 16;   // 16-byte chunks per row
    const uint8_t* gres = reinterpret_cast<const uint8_t*>(p.res);
    for (int id = tid; id < BM * cpr; id += CONV_THREADS) {
      const int row = id / cpr, j = id - row * cpr;
      const bool v = m0 + row < p.M;
      const uint8_t* src = v ? gres + ((size_t)(m0 + row) * p.Cout + n0) * res_es + j * 16 : gres;
      cp_async_16(smem_u32(sRes + row * res_pitch + j * 16), src, v ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const bool y_in_place = res_es != 0 && p.y_bits == res_es * 8;   // new residual stream overwrites the tile in smem

  // ------------------------------------------------------------------------------------------------ epilogue
  uint8_t* sOut = smem + S::OUT_OFF;
  const bool stage_low = (p.mode == HAWQ_EPI_REQUANT && p.out_bits <= 8) || (p.mode == HAWQ_EPI_RESIDUAL && p.low_bits != 0);
  const int stage_bits = (p.mode == HAWQ_EPI_REQUANT) ? p.out_bits : p.low_bits;

  constexpr double kMagic = 6755399441055744.0;      // 1.5 * 2^52
  constexpr double kOffS = 4503601774854144.0;       // 2^52 + 2^31 (signed int -> double)
  constexpr double kOffU = 4503599627370496.0;       // 2^52        (non-negative int -> double)
  bool fast_done = false;

  if constexpr (EPI == EPI_FAST_LOW) {
    if (!use_slow) {
      fast_done = true;
      // clamp(RHE((acc + bias) * M)), ReLU folded into the lower clamp bound (RHE is monotone, RHE(0) = 0)
      const int lo = p.relu ? max(p.lo, 0) : p.lo, hi = p.hi;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        const int col = wn * WNT + ni * 8 + 2 * t;
        const double2 Cb = *reinterpret_cast<const double2*>(&sCb[col]);
        const double2 M = *reinterpret_cast<const double2*>(&sM[col]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int row = wm * 32 + mi * 16 + hf * 8 + g;
            const double d0 = __hiloint2double(0x43300000, acc[mi][ni][hf * 2 + 0] ^ 0x80000000) - Cb.x;
            const double d1 = __hiloint2double(0x43300000, acc[mi][ni][hf * 2 + 1] ^ 0x80000000) - Cb.y;
            const int q0 = clampi(__double2loint(__fma_rn(d0, M.x, kMagic)), lo, hi);
            const int q1 = clampi(__double2loint(__fma_rn(d1, M.y, kMagic)), lo, hi);
            *reinterpret_cast<uint16_t*>(sOut + row * S::OUT_PITCH + col) = (uint16_t)__byte_perm(q0, q1, 0x0040);
          }
        }
      }
    }
  }

  if constexpr (EPI == EPI_FAST_RES) {
    if (!use_slow) {
      fast_done = true;
      const double res_M = dyadic_to_double(p.res_m, p.res_e), low_M = dyadic_to_double(p.low_m, p.low_e);
      const int relu_floor = p.relu ? 0 : (int)0x80000000;
      int ymax = 0;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        const int col = wn * WNT + ni * 8 + 2 * t;
        const double2 Cb = *reinterpret_cast<const double2*>(&sCb[col]);
        const double2 M = *reinterpret_cast<const double2*>(&sM[col]);
        double2 M1 = make_double2(res_M, res_M);
        if (p.res_kind == 1) M1 = *reinterpret_cast<const double2*>(&sM1[col]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int row = wm * 32 + mi * 16 + hf * 8 + g;
            const double d0 = __hiloint2double(0x43300000, acc[mi][ni][hf * 2 + 0] ^ 0x80000000) - Cb.x;
            const double d1 = __hiloint2double(0x43300000, acc[mi][ni][hf * 2 + 1] ^ 0x80000000) - Cb.y;
            const int v0 = __double2loint(__fma_rn(d0, M.x, kMagic));
            const int v1 = __double2loint(__fma_rn(d1, M.y, kMagic));
            uint8_t* rptr = sRes + row * res_pitch + col * res_es;
            double r0, r1;
            if (res_es == 2) {   // uint16 residual stream: non-negative, no sign fix-up
              const uint32_t pr = *reinterpret_cast<const uint32_t*>(rptr);
              r0 = __hiloint2double(0x43300000, (int)(pr & 0xFFFFu)) - kOffU;
              r1 = __hiloint2double(0x43300000, (int)(pr >> 16)) - kOffU;
            } else {
              const int2 pr = *reinterpret_cast<const int2*>(rptr);
              r0 = __hiloint2double(0x43300000, pr.x ^ 0x80000000) - kOffS;
              r1 = __hiloint2double(0x43300000, pr.y ^ 0x80000000) - kOffS;
            }
            int y0 = max(sat_add(__double2loint(__fma_rn(r0, M1.x, kMagic)), v0), relu_floor);
            int y1 = max(sat_add(__double2loint(__fma_rn(r1, M1.y, kMagic)), v1), relu_floor);
            if (p.y_bits == 16) {
              ymax = max(ymax, max(y0, y1));
              const uint32_t packed = (uint32_t)min(y0, 65535) | ((uint32_t)min(y1, 65535) << 16);
              if (y_in_place) *reinterpret_cast<uint32_t*>(rptr) = packed;
              else if (m0 + row < p.M)
                *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + (size_t)(m0 + row) * p.Cout + n0 + col) = packed;
            } else if (p.y_bits == 32) {
              if (y_in_place) *reinterpret_cast<int2*>(rptr) = make_int2(y0, y1);
              else if (m0 + row < p.M)
                *reinterpret_cast<int2*>(reinterpret_cast<int32_t*>(p.out) + (size_t)(m0 + row) * p.Cout + n0 + col) = make_int2(y0, y1);
            }
            if (p.low_bits != 0) {
              const double l0 = __hiloint2double(0x43300000, y0 ^ 0x80000000) - kOffS;
              const double l1 = __hiloint2double(0x43300000, y1 ^ 0x80000000) - kOffS;
              const int q0 = clampi(__double2loint(__fma_rn(l0, low_M, kMagic)), p.low_lo, p.low_hi);
              const int q1 = clampi(__double2loint(__fma_rn(l1, low_M, kMagic)), p.low_lo, p.low_hi);
              *reinterpret_cast<uint16_t*>(sOut + row * S::OUT_PITCH + col) = (uint16_t)__byte_perm(q0, q1, 0x0040);
            }
          }
        }
      }
      if (p.y_bits == 16 && ymax > 65535) atomicOr(p.status, HAWQ_FLAG_RESIDUAL_OVERFLOW);
    }
  }

  auto epilogue = [&](auto fast_tag) {
    constexpr bool FAST = decltype(fast_tag)::value;
    const double res_M = dyadic_to_double(p.res_m, p.res_e), low_M = dyadic_to_double(p.low_m, p.low_e);
    auto rq = [&](int32_t v, uint32_t m, int e, double M) -> int32_t {
      if constexpr (FAST) return rhe_requant_fast(v, M);
      else return rhe_requant(v, m, e);
    };
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int row = wm * 32 + mi * 16 + hf * 8 + g;
        const int m = m0 + row;
        const bool ok = m < p.M;
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
          const int col = wn * WNT + ni * 8 + 2 * t;
          const int4 c0 = *reinterpret_cast<const int4*>(&sChan[col]);
          const int4 c1 = *reinterpret_cast<const int4*>(&sChan[col + 1]);
          const double2 M01 = *reinterpret_cast<const double2*>(&sM[col]);
          int32_t v0 = sat_add(acc[mi][ni][hf * 2 + 0], c0.x);
          int32_t v1 = sat_add(acc[mi][ni][hf * 2 + 1], c1.x);
          const size_t gidx = (size_t)m * p.Cout + n0 + col;
          if (p.mode == HAWQ_EPI_REQUANT) {
            if (p.relu) { v0 = max(v0, 0); v1 = max(v1, 0); }
            const int32_t q0 = clampi(rq(v0, (uint32_t)c0.y, c0.z, M01.x), p.lo, p.hi);
            const int32_t q1 = clampi(rq(v1, (uint32_t)c1.y, c1.z, M01.y), p.lo, p.hi);
            if (p.out_bits <= 8) {
              *reinterpret_cast<uint16_t*>(sOut + row * S::OUT_PITCH + col) = (uint16_t)((q0 & 0xFF) | ((q1 & 0xFF) << 8));
            } else if (ok) {
              if (p.out_bits == 16) {
                *reinterpret_cast<uint32_t*>(reinterpret_cast<int16_t*>(p.out) + gidx) =
                    (uint32_t)(q0 & 0xFFFF) | ((uint32_t)(q1 & 0xFFFF) << 16);
              } else {
                *reinterpret_cast<int2*>(reinterpret_cast<int32_t*>(p.out) + gidx) = make_int2(q0, q1);
              }
            }
          } else if (p.mode == HAWQ_EPI_RESIDUAL) {
            int32_t r0 = 0, r1 = 0;
            uint32_t rm0 = p.res_m, rm1 = p.res_m;
            int re0 = p.res_e, re1 = p.res_e;
            double rM0 = res_M, rM1 = res_M;
            uint8_t* rptr = sRes + row * res_pitch + col * res_es;
            if (res_es == 2) {
              const uint32_t pr = *reinterpret_cast<const uint32_t*>(rptr);
              r0 = (int32_t)(pr & 0xFFFFu);
              r1 = (int32_t)(pr >> 16);
            } else {
              const int2 pr = *reinterpret_cast<const int2*>(rptr);
              r0 = pr.x;
              r1 = pr.y;
            }
            if (p.res_kind == 1) {
              if constexpr (FAST) {
                const double2 M1 = *reinterpret_cast<const double2*>(&sM1[col]);
                rM0 = M1.x; rM1 = M1.y;
              } else {
                rm0 = sResChan[col].m; re0 = sResChan[col].e; rm1 = sResChan[col + 1].m; re1 = sResChan[col + 1].e;
              }
            }
            int32_t y0 = sat_add(rq(r0, rm0, re0, rM0), rq(v0, (uint32_t)c0.y, c0.z, M01.x));
            int32_t y1 = sat_add(rq(r1, rm1, re1, rM1), rq(v1, (uint32_t)c1.y, c1.z, M01.y));
            if (p.relu) { y0 = max(y0, 0); y1 = max(y1, 0); }
            if (p.y_bits == 32) {
              if (y_in_place) *reinterpret_cast<int2*>(rptr) = make_int2(y0, y1);
              else if (ok) *reinterpret_cast<int2*>(reinterpret_cast<int32_t*>(p.out) + gidx) = make_int2(y0, y1);
            } else if (p.y_bits == 16) {
              if (ok && max(y0, y1) > 65535) atomicOr(p.status, HAWQ_FLAG_RESIDUAL_OVERFLOW);
              const uint32_t packed = (uint32_t)min(y0, 65535) | ((uint32_t)min(y1, 65535) << 16);
              if (y_in_place) *reinterpret_cast<uint32_t*>(rptr) = packed;
              else if (ok) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + gidx) = packed;
            }
            if (p.low_bits != 0) {
              const int32_t q0 = clampi(rq(y0, p.low_m, p.low_e, low_M), p.low_lo, p.low_hi);
              const int32_t q1 = clampi(rq(y1, p.low_m, p.low_e, low_M), p.low_lo, p.low_hi);
              *reinterpret_cast<uint16_t*>(sOut + row * S::OUT_PITCH + col) = (uint16_t)((q0 & 0xFF) | ((q1 & 0xFF) << 8));
            }
          } else if (p.mode == HAWQ_EPI_RAW_I32) {
            if (ok) *reinterpret_cast<int2*>(reinterpret_cast<int32_t*>(p.out) + gidx) = make_int2(v0, v1);
          } else {  // HAWQ_EPI_DEQUANT_F32
            if (ok) {
              float* o = reinterpret_cast<float*>(p.out) + (size_t)m * p.cout_store;
              const int c = n0 + col;
              if (c < p.cout_store) o[c] = __fmul_rn((float)v0, p.fscale[c]);
              if (c + 1 < p.cout_store) o[c + 1] = __fmul_rn((float)v1, p.fscale[c + 1]);
            }
          }
        }
      }
    }
  };
  if (!fast_done) {
    if (use_slow) epilogue(std::false_type{});
    else epilogue(std::true_type{});
  }

  if (y_in_place || stage_low) __syncthreads();
  if (y_in_place) {   // coalesced copy-out of the new residual stream tile
    const int cpr = BN * res_es / 16;
    uint8_t* gy = reinterpret_cast<uint8_t*>(p.out);
    for (int id = tid; id < BM * cpr; id += CONV_THREADS) {
      const int row = id / cpr, j = id - row * cpr;
      if (m0 + row < p.M)
        *reinterpret_cast<int4*>(gy + ((size_t)(m0 + row) * p.Cout + n0) * res_es + j * 16) =
            *reinterpret_cast<const int4*>(sRes + row * res_pitch + j * 16);
    }
  }
  if (stage_low) {
    uint8_t* gout = reinterpret_cast<uint8_t*>(p.mode == HAWQ_EPI_REQUANT ? p.out : p.out_low);
    if (stage_bits == 8) {
      constexpr int CPR = BN / 16;
      for (int id = tid; id < BM * CPR; id += CONV_THREADS) {
        const int row = id / CPR, j = id % CPR;
        if (m0 + row < p.M) {
          const int4 v = *reinterpret_cast<const int4*>(sOut + row * S::OUT_PITCH + j * 16);
          *reinterpret_cast<int4*>(gout + (size_t)(m0 + row) * p.Cout + n0 + j * 16) = v;
        }
      }
    } else {  // 4-bit: 32 channels -> 16 packed bytes
      constexpr int CPR = BN / 32;
      for (int id = tid; id < BM * CPR; id += CONV_THREADS) {
        const int row = id / CPR, j = id % CPR;
        if (m0 + row < p.M) {
          const uint4 a = *reinterpret_cast<const uint4*>(sOut + row * S::OUT_PITCH + j * 32);
          const uint4 b = *reinterpret_cast<const uint4*>(sOut + row * S::OUT_PITCH + j * 32 + 16);
          uint4 o;
          o.x = pack_nibbles8(a.x, a.y);
          o.y = pack_nibbles8(a.z, a.w);
          o.z = pack_nibbles8(b.x, b.y);
          o.w = pack_nibbles8(b.z, b.w);
          *reinterpret_cast<uint4*>(gout + (((size_t)(m0 + row) * p.Cout + n0 + j * 32) >> 1)) = o;
        }
      }
    }
  }
}

}  // namespace hawq

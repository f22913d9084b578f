// Resize-unit tail in one kernel: y = ReLU( RHE(m1 * (conv1x1_s(x2, w2) + bias2)) + RHE(m * (conv1x1(x, w) + bias)) ), uint16 stream +
// the next unit's low-bit activation (Q_ResUnitBn.forward with resize_identity, q_resnet.py:234-236,254-258; case 1 of
// fixedpoint_fn, quant_utils.py:416-456).  Same structure as conv1x1.cuh - stationary weights (both matrices), activations in few
// large TMA boxes, 16 epilogue warps, outputs staged and stored with one TMA operation per tile - with two differences:
//   * two TMEM accumulators per buffer (4 * BN columns): the main 1x1 convolution and the identity 1x1 convolution accumulate
//     side by side and the case-1 sum combines them, so the identity branch never exists in HBM;
//   * the identity convolution may be strided (stride 2 at the head of stages 2-4): its input rows are fetched by ONE 5-D TMA
//     box {64 B, 2 Wo, 2 R, NI, KC} with traversal strides {1, 2, 2, 1, 1} over the NHWC tensor - the TMA unit does the
//     sub-sampling.  For that an output tile is R whole rows of one image (or NI whole images) = TR <= 128 rows; the k-tile
//     blocks in shared memory are then TR * 64 B apart and the MMA (M = 128) reads past the block end: rows >= TR of the
//     accumulator are garbage and never stored.
// A4: packed 4-bit activations ({32 B, ...} boxes) are expanded to int8 by four converter warps, once per stage.
#pragma once
#include "tc_ptx.cuh"

namespace hawq {

struct DualParams {
  const hawq_chan* chan;     // main conv: bias, case-1 ratio (m2, e2) per channel
  const hawq_chan* chan2;    // identity conv: bias, case-1 ratio (m1, e1) per channel
  int32_t* status;
  int M, Cout;
  int KT1, KT2, KC;          // k-tiles of the two convolutions, k-tiles per stage (divides both)
  int NS;                    // activation stages
  int TR;                    // rows per tile (128, or R * Wo * NI)
  int strided;               // identity input fetched by the strided 5-D box
  int Wo, HoWo, R, stride2;  // output geometry (strided mode): tile -> (image, first row)
  int m_tiles, n_tiles, ctas_per_n;
  int w1_boxes, w1_box_kt, w2_boxes, w2_box_kt;
  int low_bits; uint32_t low_m; int low_e, low_lo, low_hi;
  int sat_pack;
  int out_bufs, y_stride, low_stride;   // staged output tiles (2 = double-buffered), bytes between the buffers
  int off_a, off_packed, off_y, off_low, off_cst, off_bar;   // shared-memory carve-up (weights at 0: [W1 | W2])
};

constexpr int DUAL_EPI_WARPS = 16;
constexpr int DUAL_MAX_STAGES = 4;
__host__ __device__ constexpr int dual_producer_warps(bool a4) { return a4 ? 4 : 1; }
__host__ __device__ constexpr int dual_threads(bool a4) { return (dual_producer_warps(a4) + 1 + DUAL_EPI_WARPS) * 32; }

struct alignas(64) DualMaps {
  CUtensorMap a;     // main activations {64 B, M rows, KT1}: box {64, TR, KC}
  CUtensorMap a2;    // identity activations: {64 B, M rows, KT2} box {64, TR, KC}, or 5-D {64 B, W2, H2, N, KT2} box {64, 2 Wo, 2 R, NI, KC} / strides {1, s, s, 1, 1}
  CUtensorMap w1, w2;   // weights {64 B, Cout rows, KT}: box {64, BN, box_kt}
  CUtensorMap y;     // uint16 stream {128 B, M rows, 2 Cout / 128}: box {128, TR, BN / 64}, SWIZZLE_128B
  CUtensorMap low;   // low-bit output [M][Cout * bits / 8]: box {BN * bits / 8, TR}
};

__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}

template <int BN, bool WIDE, bool A4>
__global__ void __launch_bounds__(dual_threads(A4), 1) conv_dual_kernel(const DualParams p, const __grid_constant__ DualMaps maps) {
  constexpr int B_STAGE = BN * 64;
  constexpr int NPW = dual_producer_warps(A4);
  constexpr int MMA_WARP = NPW, EPI_WARP0 = NPW + 1;
  constexpr int CW = BN / 4;                 // columns per epilogue warp: 16 / 32
  constexpr int TMEM_COLS = 4 * BN;          // two accumulators x two buffers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  double2* sCst = reinterpret_cast<double2*>(smem + p.off_cst);          // main: {2^52 + 2^31 - bias, M}
  double2* sCst2 = sCst + BN;                                            // identity
  const uint32_t bar_base = smem_base + p.off_bar;
  const uint32_t b_full = bar_base;
  auto afull = [&](int s) { return bar_base + 8u * (1 + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (1 + DUAL_MAX_STAGES + s); };
  auto kfull = [&](int s) { return bar_base + 8u * (1 + 2 * DUAL_MAX_STAGES + s); };    // A4: packed stage landed
  auto kempty = [&](int s) { return bar_base + 8u * (1 + 3 * DUAL_MAX_STAGES + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (1 + 4 * DUAL_MAX_STAGES + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (3 + 4 * DUAL_MAX_STAGES + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (5 + 4 * DUAL_MAX_STAGES));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x % p.n_tiles, slot = blockIdx.x / p.n_tiles;
  const int n0 = nt * BN;
  const int my_tiles = (slot < p.m_tiles) ? (p.m_tiles - 1 - slot) / p.ctas_per_n + 1 : 0;
  const int KT = p.KT1 + p.KT2;

  if (tid == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < DUAL_MAX_STAGES; ++s) {
      mbar_init(afull(s), A4 ? 4 : 1);        // A4: one arrival per converter warp
      mbar_init(aempty(s), 1);
      mbar_init(kfull(s), 1);
      mbar_init(kempty(s), 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), DUAL_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  asm volatile("griddepcontrol.launch_dependents;");

  if (warp < NPW) {
    // =============================================================================== producer (+ A4 converters)
    if (warp == 0 && elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)KT * B_STAGE);      // plan-time data: before waiting for the previous kernel
      for (int i = 0; i < p.w1_boxes; ++i)
        tma_load_3d(smem_base + (uint32_t)(i * p.w1_box_kt) * B_STAGE, &maps.w1, 0, n0, i * p.w1_box_kt, b_full);
      for (int i = 0; i < p.w2_boxes; ++i)
        tma_load_3d(smem_base + (uint32_t)(p.KT1 + i * p.w2_box_kt) * B_STAGE, &maps.w2, 0, n0, i * p.w2_box_kt, b_full);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t stage_bytes = (uint32_t)p.KC * p.TR * (A4 ? 32 : 64);
    const uint32_t stage_alloc = (uint32_t)p.KC * 128 * 64, packed_alloc = (uint32_t)p.KC * 128 * 32;
    const int S1 = p.KT1 / p.KC, S2 = p.KT2 / p.KC;               // stages per tile: main conv, identity conv
    // stage j of tile t: j < S1 -> main conv k-tiles j * KC ..., else identity conv k-tiles (j - S1) * KC ...
    auto load_stage = [&](int t, int j, uint32_t dst, uint32_t bar) {
      const int m0 = (slot + t * p.ctas_per_n) * p.TR;
      if (j < S1) {
        tma_load_3d(dst, &maps.a, 0, m0, j * p.KC, bar);
      } else if (p.strided) {
        const int img0 = m0 / p.HoWo, y0 = (m0 - img0 * p.HoWo) / p.Wo;
        tma_load_5d(dst, &maps.a2, 0, 0, y0 * p.stride2, img0, (j - S1) * p.KC, bar);
      } else {
        tma_load_3d(dst, &maps.a2, 0, m0, (j - S1) * p.KC, bar);
      }
    };
    if constexpr (!A4) {
      if (elect_one()) {
        uint32_t s = 0, ph = 0;
        for (int t = 0; t < my_tiles; ++t)
          for (int j = 0; j < S1 + S2; ++j) {
            mbar_wait_small(aempty(s), ph ^ 1);
            mbar_arrive_expect_tx(afull(s), stage_bytes);
            load_stage(t, j, smem_base + p.off_a + s * stage_alloc, afull(s));
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1; }
          }
      }
    } else {
      // packed 4-bit rows: TMA -> packed stage ([k-tile][TR rows][32 B], SWIZZLE_32B) -> these 128 threads expand every row to int8 in
      // the K order the permuted weights expect -> [k-tile][TR rows][64 B], SWIZZLE_64B (swizzle phase = absolute row kt * TR + r)
      const int spt = S1 + S2, total_g = my_tiles * spt;
      auto issue = [&](int g) {            // one elected lane of warp 0
        const int t = g / spt, j = g - t * spt, ks = g % p.NS;
        mbar_wait_small(kempty(ks), ((g / p.NS) & 1) ^ 1);
        mbar_arrive_expect_tx(kfull(ks), stage_bytes);
        load_stage(t, j, smem_base + p.off_packed + ks * packed_alloc, kfull(ks));
      };
      if (warp == 0) {
        for (int g = 0; g < p.NS - 1 && g < total_g; ++g)
          if (elect_one()) issue(g);
        __syncwarp();
      }
      for (int g = 0; g < total_g; ++g) {
        if (warp == 0) {
          if (g + p.NS - 1 < total_g && elect_one()) issue(g + p.NS - 1);
          __syncwarp();
        }
        const int s = g % p.NS;
        mbar_wait_small(kfull(s), (g / p.NS) & 1);
        mbar_wait_small(aempty(s), ((g / p.NS) & 1) ^ 1);
        const uint8_t* src = smem + p.off_packed + s * packed_alloc;
        uint8_t* dst = smem + p.off_a + s * stage_alloc;
        if (tid < p.TR) {
          uint4 wv[4][2];
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
            if (kt < p.KC) {
              const int pr = kt * p.TR + tid;
#pragma unroll
              for (int blk = 0; blk < 2; ++blk) wv[kt][blk] = *reinterpret_cast<const uint4*>(src + pr * 32 + ((blk ^ ((pr >> 2) & 1)) << 4));
            }
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
            if (kt < p.KC) {
              const int pr = kt * p.TR + tid;
              const uint32_t a_sw = (pr >> 1) & 3;
#pragma unroll
              for (int blk = 0; blk < 2; ++blk) {
                const uint4 v = wv[kt][blk];
                const uint4 lo = make_uint4(v.x & 0x0F0F0F0Fu, v.y & 0x0F0F0F0Fu, v.z & 0x0F0F0F0Fu, v.w & 0x0F0F0F0Fu);
                const uint4 hi = make_uint4((v.x >> 4) & 0x0F0F0F0Fu, (v.y >> 4) & 0x0F0F0F0Fu, (v.z >> 4) & 0x0F0F0F0Fu, (v.w >> 4) & 0x0F0F0F0Fu);
                *reinterpret_cast<uint4*>(dst + pr * 64 + (((2 * blk) ^ a_sw) << 4)) = lo;
                *reinterpret_cast<uint4*>(dst + pr * 64 + (((2 * blk + 1) ^ a_sw) << 4)) = hi;
              }
            }
        }
        fence_proxy_async();             // generic-proxy writes -> tcgen05.mma (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(afull(s));
          mbar_arrive(kempty(s));
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================================================================== MMA issuer
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_i8(128, BN, !A4);          // packed 4-bit activations are unsigned
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw64(0) >> 32);
      constexpr uint32_t BU = B_STAGE >> 4;
      const uint32_t AU = (uint32_t)p.TR * 4u;                    // one activation k-tile (TR rows x 64 B) in descriptor units
      const uint32_t a_base = ((smem_base + p.off_a) >> 4) | (1u << 16), a_step = (uint32_t)p.KC * 512u;
      const uint32_t w_base = (smem_base >> 4) | (1u << 16);
      mbar_wait_small(b_full, 0);
      uint32_t s = 0, ph = 0, a0 = a_base;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        mbar_wait_small(tempty(buf), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        uint32_t wl = w_base;
#pragma unroll 1
        for (int conv = 0; conv < 2; ++conv) {                    // main conv -> accumulator 0, identity conv -> accumulator 1
          const uint32_t d_tmem = tmem_base + buf * (2 * BN) + conv * BN;
          const int kt_conv = conv ? p.KT2 : p.KT1;
          for (int k0 = 0; k0 < kt_conv; k0 += p.KC) {
            mbar_wait_small(afull(s), ph);
            tc_fence_after();
            uint32_t al = a0;
            if (k0 == 0) umma_i8_lohi<false>(d_tmem, al, wl, desc_hi, idesc);
            else umma_i8_lohi<true>(d_tmem, al, wl, desc_hi, idesc);
            umma_i8_lohi<true>(d_tmem, al + 2, wl + 2, desc_hi, idesc);
            for (int kt = 1; kt < p.KC; ++kt) {
              al += AU; wl += BU;
              umma_i8_lohi<true>(d_tmem, al, wl, desc_hi, idesc);
              umma_i8_lohi<true>(d_tmem, al + 2, wl + 2, desc_hi, idesc);
            }
            wl += BU;
            umma_commit(aempty(s));
            a0 += a_step;
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1; a0 = a_base; }
          }
        }
        umma_commit(tfull(buf));
      }
    }
  } else {
    // =============================================================================== epilogue (16 warps)
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                      // column group of CW columns
    constexpr double kMagic = 6755399441055744.0, kOffS = 4503601774854144.0, kOffU = 4503599627370496.0;
    int bad = 0, ovf = 0, ymax = 0;
    auto ratio_ok = [](uint32_t m_, int e_) { return m_ == 0u || e_ >= (WIDE ? 11 : 31); };
    for (int i = tid - EPI_WARP0 * 32; i < BN; i += DUAL_EPI_WARPS * 32) {     // plan-time data
      const hawq_chan ch = p.chan[n0 + i], c2 = p.chan2[n0 + i];
      sCst[i] = make_double2(kOffS - (double)ch.bias, dyadic_to_double(ch.m, ch.e));
      sCst2[i] = make_double2(kOffS - (double)c2.bias, dyadic_to_double(c2.m, c2.e));
      bad |= !ratio_ok(ch.m, ch.e) | (ch.bias >= (1 << 29)) | (ch.bias <= -(1 << 29));
      bad |= !ratio_ok(c2.m, c2.e) | (c2.bias >= (1 << 29)) | (c2.bias <= -(1 << 29));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DUAL_EPI_WARPS * 32));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int row = quarter * 32 + lane;
    const bool valid_row = row < p.TR;           // rows >= TR of the accumulator are garbage (the MMA read past the k-tile block)
    const double2* cst = sCst + cg * CW;
    const double2* cst2 = sCst2 + cg * CW;
    const bool elect_x = (ew == 0 && lane == 0);     // issues the TMA stores
    const double low_M = dyadic_to_double(p.low_m, p.low_e);
    const double low_C = kMagic - kOffU * low_M;
    const bool sat8 = p.sat_pack != 0 && p.low_bits == 8 && p.low_hi == 127 && p.low_lo <= 0;
    // ... and with clamp [<= 0, hi <= 255] (4-bit values in byte containers) the u8 saturation followed by a per-byte min
    const bool satu = p.sat_pack != 0 && p.low_bits == 8 && !sat8 && p.low_lo <= 0 && p.low_hi >= 0 && p.low_hi <= 255;
    const uint32_t hi4 = (uint32_t)(p.low_hi & 255) * 0x01010101u;
    const int l_lo = p.low_lo, l_hi = p.low_hi;
    if (p.low_bits) bad |= !dyadic_is_fast(p.low_m, p.low_e) | (p.low_m != 0u && p.low_e > 51);
    const uint32_t r_chunk = (uint32_t)(cg * CW) / 64, r_piece0 = ((uint32_t)(cg * CW) % 64) / 8;
    for (int t = 0; t < my_tiles; ++t) {
      const uint32_t buf = t & 1;
      mbar_wait_small(tfull(buf), (t >> 1) & 1);
      tc_fence_after();
      uint4 yo[CW / 8];
      uint32_t lw[CW / 4];
#pragma unroll
      for (int h = 0; h < CW / 16; ++h) {          // 16 columns at a time: both accumulators
        uint32_t acc[16], acc2[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (2 * BN) + cg * CW + h * 16;
        tmem_ld16(taddr, acc);
        tmem_ld16(taddr + BN, acc2);
        tmem_ld_wait();
        if (h == CW / 16 - 1) {                    // everything is in registers: hand the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {              // groups of 8 channels
          int y[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = h * 16 + i * 8 + k;
            const double2 cm = cst[c], c2 = cst2[c];
            const double qv = __fma_rn(__hiloint2double(0x43300000, acc[i * 8 + k] ^ 0x80000000) - cm.x, cm.y, kMagic);
            const double qr = __fma_rn(__hiloint2double(0x43300000, acc2[i * 8 + k] ^ 0x80000000) - c2.x, c2.y, kMagic);
            const int v = __double2loint(qv), vr = __double2loint(qr);
            const int sum = v + vr;
            if constexpr (WIDE) {
              int o = (__double2hiint(qv) + (int)((uint32_t)v >> 31)) ^ 0x43380000;
              o |= (__double2hiint(qr) + (int)((uint32_t)vr >> 31)) ^ 0x43380000;
              o |= ((v ^ sum) & (vr ^ sum)) >> 31;                     // the sum itself wrapped
              ovf |= valid_row ? o : 0;
            }
            y[k] = max(sum, 0);
            ymax = max(ymax, valid_row ? y[k] : 0);
          }
          const int g8 = h * 2 + i;
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[g8].x) : "r"(y[1]), "r"(y[0]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[g8].y) : "r"(y[3]), "r"(y[2]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[g8].z) : "r"(y[5]), "r"(y[4]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[g8].w) : "r"(y[7]), "r"(y[6]));
          if (p.low_bits) {
            int q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = __double2loint(__fma_rn(__hiloint2double(0x43300000, y[k]), low_M, low_C));   // y >= 0
            if (sat8) {
              uint32_t h0, h1;
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(h0) : "r"(q[3]), "r"(q[2]), "r"(0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lw[2 * g8]) : "r"(q[1]), "r"(q[0]), "r"(h0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(h1) : "r"(q[7]), "r"(q[6]), "r"(0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lw[2 * g8 + 1]) : "r"(q[5]), "r"(q[4]), "r"(h1));
            } else if (satu) {
              uint32_t h0, h1, o0, o1;
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(h0) : "r"(q[3]), "r"(q[2]), "r"(0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(o0) : "r"(q[1]), "r"(q[0]), "r"(h0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(h1) : "r"(q[7]), "r"(q[6]), "r"(0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(o1) : "r"(q[5]), "r"(q[4]), "r"(h1));
              lw[2 * g8] = __vminu4(o0, hi4);
              lw[2 * g8 + 1] = __vminu4(o1, hi4);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) q[k] = clampi(q[k], l_lo, l_hi);
              lw[2 * g8] = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
              lw[2 * g8 + 1] = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
            }
          }
        }
      }
      // stage y ([64-column chunk][TR rows][128 B], SWIZZLE_128B) and the low-bit tile; one TMA store each per tile
      const uint32_t y_off = p.off_y + (p.out_bufs == 2 ? (t & 1) * p.y_stride : 0);
      const uint32_t low_off = p.off_low + (p.out_bufs == 2 ? (t & 1) * p.low_stride : 0);
      if (elect_x) { if (p.out_bufs == 2) bulk_wait_read_1(); else bulk_wait_read_all(); }   // the stores that last read these staging tiles are done
      asm volatile("bar.sync 1, %0;" ::"n"(DUAL_EPI_WARPS * 32));
      if (valid_row) {
        // the staged tile is [64-column chunk][TR rows][128 B]; the swizzle is a function of the shared-memory address, and a
        // chunk of TR rows need not start on a 1024-byte boundary: index it as one tile of (chunk * TR + row) rows
        const int yrow = (int)r_chunk * p.TR + row;
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) *reinterpret_cast<uint4*>(smem + y_off + tile_piece_off(128, yrow, (int)r_piece0 + i)) = yo[i];
        if (p.low_bits) {
          uint8_t* lt = smem + low_off;
          const int rb_low = BN * p.low_bits / 8;                 // 128 / 64 / 32
          if (p.low_bits == 8) {
#pragma unroll
            for (int j = 0; j < CW / 16; ++j)
              *reinterpret_cast<uint4*>(lt + tile_piece_off(rb_low, row, cg * (CW / 16) + j)) = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < CW / 16; ++j) {
              const int boff = cg * (CW / 2) + j * 8;
              *reinterpret_cast<uint2*>(lt + tile_piece_off(rb_low, row, boff >> 4) + (boff & 8)) =
                  make_uint2(pack_nibbles8(lw[4 * j], lw[4 * j + 1]), pack_nibbles8(lw[4 * j + 2], lw[4 * j + 3]));
            }
          }
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(DUAL_EPI_WARPS * 32));
      if (elect_x) {
        const int m0 = (slot + t * p.ctas_per_n) * p.TR;
        tma_store_3d(&maps.y, 0, m0, n0 / 64, smem_base + y_off);
        if (p.low_bits) tma_store_2d(&maps.low, n0 * p.low_bits / 8, m0, smem_base + low_off);
        bulk_commit();
      }
    }
    if (elect_x) bulk_wait_all();
    if (ymax > 65535) atomicOr(p.status, HAWQ_FLAG_RESIDUAL_OVERFLOW);
    if (bad) atomicOr(p.status, HAWQ_FLAG_BAD_RATIO);
    if (ovf) atomicOr(p.status, HAWQ_FLAG_REQUANT_OVERFLOW);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace hawq

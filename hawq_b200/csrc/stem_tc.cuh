// Stem of the quantized ResNets in ONE kernel on 5th-gen tensor cores: 7x7 stride-2 pad-3 convolution (Cin = 3 -> 64) + bias,
// nn.MaxPool2d(3, 2, 1), the 16-bit dyadic requantisation of quant_act_int32 (clamped), ReLU, and the first unit's low-bit
// quant_act - reference utils/models/q_resnet.py:117-122 (+ :234 for the low-bit copy).  Requantisation and ReLU are monotone
// (positive multiplier), so they commute with the max-pool: every convolution output is requantised once and the pool takes the
// maximum of int16 values; the int16 convolution output never reaches HBM.
//
// Work unit = (image, band of PB pooled rows).  Per unit:
//   raw      one 3-D TMA box {W * 3 / 4 words, 4 PB + 7 input rows, 1 image} (rows outside the image zero-filled) -> raw bytes
//   pix      8 builder warps expand the 3-byte pixels to one 32-bit word per pixel (channels 0-2 + zero), 3 zero pixels left, 5 right
//   A tile   per convolution row (112 pixels = MMA rows, 16 idle): with the K order (kh, kw padded to 8, c padded to 4) the K = 32
//            slice of pixel ox for kernel row kh is the 32 contiguous bytes pix[2 * oy + kh][2 * ox .. 2 * ox + 7]: the builders
//            copy them (4 x LDS.64 -> 2 x STS.128) into a [4 k-tiles][128][64 B] SWIZZLE_64B tile, two kernel rows per k-tile
//   MMA      one elected lane: 7 x tcgen05.mma kind::i8 (M = 128, N = 64, K = 32) per convolution row, weights [64][256] stationary
//   epilogue 16 warps: tcgen05.ld, + bias, exact FP64-FMA requantisation, clamp, ReLU -> int16 row in a 4-slot row ring; after
//            rows 2p - 1, 2p, 2p + 1 the pooled row p = 3x3 maximum -> uint16 / int32 residual stream + low-bit copy, staged and
//            stored with one TMA operation per tensor and pooled row.
// Fast-path preconditions (host-checked, else the two-kernel path of stem.cuh): every ratio <= 1 (e >= 31), low-bit e <= 51.
#pragma once
#include "tc_ptx.cuh"

namespace hawq {

struct StemParams {
  const hawq_chan* chan;
  int32_t* status;
  int N, H, W;               // input image (NHWC, 3 channels, int8)
  int Hc, Wc;                // convolution output
  int Hp, Wp;                // pooled output
  int PB;                    // pooled rows per unit
  int bands;                 // units per image
  int units;                 // N * bands
  int raw_rows;              // 4 * PB + 7
  int raw_pitch;             // W * 3 bytes
  int pix_pitch;             // words per pixel row: W + 8
  int lo, hi;                // 16-bit clamp of quant_act_int32
  int y_bits;                // 16 | 32
  int low_bits; uint32_t low_m; int low_e, low_lo, low_hi;
  int out_bufs, y_stride, low_stride;   // staged output tiles (2 = double-buffered, alternating per emitted pooled row), bytes between the buffers
  int off_raw, off_pix, off_a, off_rows, off_y, off_low, off_cst, off_bar;   // shared-memory carve-up (weights at 0)
};

struct alignas(64) StemMaps {
  CUtensorMap x;     // input as uint32 words {W * 3 / 4, H, N}: box {W * 3 / 4, raw_rows, 1}
  CUtensorMap w;     // weights [64][256] int8 as {64 B, 64 rows, 4 k-tiles}: box {64, 64, 4}, SWIZZLE_64B
  CUtensorMap y;     // pooled stream [N * Hp * Wp][64 * y_bits / 8]: box {128 B, Wp rows (, 2 chunks for int32)}, SWIZZLE_128B
  CUtensorMap low;   // low-bit copy [N * Hp * Wp][64 * low_bits / 8]: box {64 | 32 B, Wp}
};

constexpr int STEM_BUILD_WARPS = 8, STEM_EPI_WARPS = 16;
constexpr int STEM_MMA_WARP = STEM_BUILD_WARPS;
constexpr int STEM_EPI_WARP0 = STEM_BUILD_WARPS + 1;
constexpr int STEM_TC_THREADS = (STEM_BUILD_WARPS + 1 + STEM_EPI_WARPS) * 32;   // 800
constexpr int STEM_A_TILE = 4 * 128 * 64;      // one convolution row: 4 k-tiles
constexpr int STEM_ROW_BYTES = 128 * 128;      // one requantised convolution row: <= 128 pixels x 64 channels x int16 (the last slot may be shorter: Wc pixels)

__global__ void __launch_bounds__(STEM_TC_THREADS, 1) stem_tc_kernel(const StemParams p, const __grid_constant__ StemMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  double2* sCst = reinterpret_cast<double2*>(smem + p.off_cst);
  const uint32_t bar_base = smem_base + p.off_bar;
  const uint32_t b_full = bar_base;
  const uint32_t raw_full = bar_base + 8, raw_empty = bar_base + 16;
  auto afull = [&](int s) { return bar_base + 8u * (3 + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (5 + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (7 + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (9 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_units = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // convolution rows of unit u (band b of an image): first = max(2 * p0 - 1, 0), last = min(2 * (p0 + PB - 1) + 1, Hc - 1)
  auto unit_rows = [&](int u, int& n_img, int& p0, int& c_first, int& c_last) {
    const int unit = blockIdx.x + u * gridDim.x;
    n_img = unit / p.bands;
    p0 = (unit - n_img * p.bands) * p.PB;
    c_first = max(2 * p0 - 1, 0);
    c_last = min(2 * (min(p0 + p.PB, p.Hp) - 1) + 1, p.Hc - 1);
  };

  if (tid == 0) {
    mbar_init(b_full, 1);
    mbar_init(raw_full, 1);
    mbar_init(raw_empty, STEM_BUILD_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(afull(s), STEM_BUILD_WARPS);
      mbar_init(aempty(s), 1);
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), STEM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == STEM_MMA_WARP) tmem_alloc<128>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < STEM_BUILD_WARPS) {
    // =============================================================================== builders (256 threads)
    if (warp == 0 && elect_one()) {               // weights: plan-time data
      mbar_arrive_expect_tx(b_full, 64 * 256);
      tma_load_3d(smem_base, &maps.w, 0, 0, 0, b_full);
    }
    // zero the side pads of the pixel rows once (3 pixels left, 5 right)
    uint32_t* pix = reinterpret_cast<uint32_t*>(smem + p.off_pix);
    for (int i = tid; i < p.raw_rows * 8; i += STEM_BUILD_WARPS * 32) {
      const int r = i >> 3, j = i & 7;
      pix[r * p.pix_pitch + (j < 3 ? j : p.W + j)] = 0u;
    }
    const uint32_t raw_bytes = (uint32_t)p.raw_rows * p.raw_pitch;
    auto issue_raw = [&](int u) {                 // one elected lane of warp 0
      int n_img, p0, c_first, c_last;
      unit_rows(u, n_img, p0, c_first, c_last);
      mbar_wait_small(raw_empty, (u & 1) ^ 1);
      mbar_arrive_expect_tx(raw_full, raw_bytes);
      tma_load_3d(smem_base + p.off_raw, &maps.x, 0, 2 * c_first - 3, n_img, raw_full);
    };
    if (warp == 0) {
      if (my_units > 0 && elect_one()) issue_raw(0);
      __syncwarp();
    }
    uint32_t g = 0;                               // convolution rows built so far (A ring position)
    for (int u = 0; u < my_units; ++u) {
      int n_img, p0, c_first, c_last;
      unit_rows(u, n_img, p0, c_first, c_last);
      // ---- expand the raw rows: 4 pixels (12 bytes) -> 4 words
      mbar_wait_small(raw_full, u & 1);
      const int groups = p.W >> 2;                // 4-pixel groups per row
      for (int i = tid; i < p.raw_rows * groups; i += STEM_BUILD_WARPS * 32) {
        const int r = i / groups, q = i - r * groups;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(smem + p.off_raw + r * p.raw_pitch) + 3 * q;
        const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
        uint4 o;
        o.x = w0 & 0x00FFFFFFu;
        o.y = __byte_perm(w0, w1, 0x4543) & 0x00FFFFFFu;     // bytes 3 (w0), 4, 5 (w1)
        o.z = __byte_perm(w1, w2, 0x4432) & 0x00FFFFFFu;     // bytes 2, 3 (w1), 4 (w2)
        o.w = w2 >> 8;
        uint32_t* dstp = pix + r * p.pix_pitch + 3 + 4 * q;                  // (3 pad words: the builders' 8-word windows start on even words)
        dstp[0] = o.x; dstp[1] = o.y; dstp[2] = o.z; dstp[3] = o.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(raw_empty);       // this warp is done with the raw buffer
      asm volatile("bar.sync 2, %0;" ::"n"(STEM_BUILD_WARPS * 32));
      if (warp == 0) {                             // prefetch the next unit's rows while this unit's tiles are built
        if (u + 1 < my_units && elect_one()) issue_raw(u + 1);
        __syncwarp();
      }
      // ---- one A tile per convolution row
      const int r = tid & 127, par = tid >> 7;     // MMA row (output pixel), kernel rows par, par + 2, ...
      for (int c = c_first; c <= c_last; ++c, ++g) {
        const uint32_t s = g & 1;
        mbar_wait_small(aempty(s), ((g >> 1) & 1) ^ 1);
        if (r < p.Wc) {
          uint8_t* at = smem + p.off_a + s * STEM_A_TILE + r * 64;
          const uint32_t sw = (r >> 1) & 3;
          // input row of (convolution row c, kernel row kh) = 2 c - 3 + kh = raw row 2 (c - c_first) + kh; pixel 2 r - 3 + kw = word 2 r + kw
          const uint32_t* prow = pix + 2 * (c - c_first) * p.pix_pitch + 2 * r;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int kh = par + 2 * kk;
            if (kh < 7) {
              const uint2* sp = reinterpret_cast<const uint2*>(prow + kh * p.pix_pitch);
              const uint2 a = sp[0], b = sp[1], cc = sp[2], d = sp[3];
              uint8_t* dst = at + (kh >> 1) * (128 * 64);
              const int piece = (kh & 1) * 2;
              *reinterpret_cast<uint4*>(dst + (((piece) ^ sw) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
              *reinterpret_cast<uint4*>(dst + (((piece + 1) ^ sw) << 4)) = make_uint4(cc.x, cc.y, d.x, d.y);
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull(s));
      }
      asm volatile("bar.sync 2, %0;" ::"n"(STEM_BUILD_WARPS * 32));   // the pixel rows are free for the next unit
    }
  } else if (warp == STEM_MMA_WARP) {
    // =============================================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_i8(128, 64, true);
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw64(0) >> 32);
      const uint32_t w0 = (smem_base >> 4) | (1u << 16);
      const uint32_t a_base = ((smem_base + p.off_a) >> 4) | (1u << 16);
      mbar_wait_small(b_full, 0);
      uint32_t g = 0;
      for (int u = 0; u < my_units; ++u) {
        int n_img, p0, c_first, c_last;
        unit_rows(u, n_img, p0, c_first, c_last);
        for (int c = c_first; c <= c_last; ++c, ++g) {
          const uint32_t s = g & 1;
          mbar_wait_small(tempty(s), ((g >> 1) & 1) ^ 1);
          mbar_wait_small(afull(s), (g >> 1) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + s * 64;
          const uint32_t a0 = a_base + s * (STEM_A_TILE >> 4);
          umma_i8_lohi<false>(d_tmem, a0, w0, desc_hi, idesc);
#pragma unroll
          for (int kh = 1; kh < 7; ++kh)
            umma_i8_lohi<true>(d_tmem, a0 + (kh >> 1) * 512 + (kh & 1) * 2, w0 + (kh >> 1) * 256 + (kh & 1) * 2, desc_hi, idesc);
          umma_commit(aempty(s));
          umma_commit(tfull(s));
        }
      }
    }
  } else {
    // =============================================================================== epilogue (16 warps)
    const int ew = warp - STEM_EPI_WARP0;
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                      // 16 channels
    const int etid = tid - STEM_EPI_WARP0 * 32;  // 0 .. 511
    constexpr double kMagic = 6755399441055744.0, kOffS = 4503601774854144.0, kOffU = 4503599627370496.0;
    int bad = 0;
    for (int i = etid; i < 64; i += STEM_EPI_WARPS * 32) {
      const hawq_chan ch = p.chan[i];
      sCst[i] = make_double2(kOffS - (double)ch.bias, dyadic_to_double(ch.m, ch.e));
      bad |= !(ch.m == 0u || ch.e >= 31) | (ch.bias >= (1 << 29)) | (ch.bias <= -(1 << 29));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(STEM_EPI_WARPS * 32));
    const double low_M = dyadic_to_double(p.low_m, p.low_e);
    const double low_C = kMagic - kOffU * low_M;
    if (p.low_bits) bad |= !dyadic_is_fast(p.low_m, p.low_e) | (p.low_m != 0u && p.low_e > 51);
    const int px = quarter * 32 + lane;          // convolution pixel of this thread
    const double2* cst = sCst + cg * 16;
    const bool elect_x = etid == 0;
    const int q_lo = max(p.lo, 0), q_hi = p.hi;  // clamp, then ReLU
    // pooling role of this thread: pooled pixel pp (8 threads per pixel), channels 8 * pc .. 8 * pc + 7
    const int pp = etid >> 3, pc = etid & 7;
    uint32_t g = 0, emitted = 0;
    for (int u = 0; u < my_units; ++u) {
      int n_img, p0, c_first, c_last;
      unit_rows(u, n_img, p0, c_first, c_last);
      for (int c = c_first; c <= c_last; ++c, ++g) {
        const uint32_t s = g & 1;
        mbar_wait_small(tfull(s), (g >> 1) & 1);
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 64 + cg * 16, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(s));
        uint32_t w16[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          int q[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const double2 cm = cst[j + k];
            const double d = __hiloint2double(0x43300000, acc[j + k] ^ 0x80000000) - cm.x;
            q[k] = clampi(__double2loint(__fma_rn(d, cm.y, kMagic)), q_lo, q_hi);
          }
          w16[j >> 1] = (uint32_t)q[0] | ((uint32_t)q[1] << 16);     // 0 <= q <= 32767
        }
        // requantised row -> row ring slot c & 3: [pixel][64 channels] int16, 16-byte pieces XOR-swizzled by the pixel index
        uint8_t* rowp = smem + p.off_rows + (c & 3) * STEM_ROW_BYTES + px * 128;
        if (px < p.Wc) {                            // (the last ring slot only holds Wc pixels)
          *reinterpret_cast<uint4*>(rowp + (((2 * cg) ^ (px & 7)) << 4)) = make_uint4(w16[0], w16[1], w16[2], w16[3]);
          *reinterpret_cast<uint4*>(rowp + (((2 * cg + 1) ^ (px & 7)) << 4)) = make_uint4(w16[4], w16[5], w16[6], w16[7]);
        }
        // a pooled row is complete after convolution row 2 * pr + 1 (or the last row of the image)
        const bool emit = (c & 1) || c == p.Hc - 1;
        if (!emit) continue;
        const int pr = c >> 1;
        if (pr < p0 || pr >= p0 + p.PB || pr >= p.Hp) continue;          // (row 2 * p0 - 1 only feeds the first pooled row of the band)
        const uint32_t ob = p.out_bufs == 2 ? (emitted++ & 1) : 0;         // staging buffers alternate per emitted pooled row
        const uint32_t y_off = p.off_y + ob * p.y_stride;
        const uint32_t low_off = p.off_low + ob * p.low_stride;
        if (elect_x) { if (p.out_bufs == 2) bulk_wait_read_1(); else bulk_wait_read_all(); }   // the stores that last read these staging tiles are done
        asm volatile("bar.sync 1, %0;" ::"n"(STEM_EPI_WARPS * 32));      // the three rows are complete, the staging tiles are free
        if (pp < p.Wp) {
          // rows 2 pr - 1, 2 pr, 2 pr + 1 and pixels 2 pp - 1, 2 pp, 2 pp + 1, clipped by duplication (all values are >= 0)
          const int r0 = max(2 * pr - 1, 0), r2 = min(2 * pr + 1, p.Hc - 1);
          const int x0 = max(2 * pp - 1, 0), x2 = min(2 * pp + 1, p.Wc - 1);
          uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int rr = 0; rr < 3; ++rr) {
            const int row = rr == 0 ? r0 : rr == 1 ? 2 * pr : r2;
            const uint8_t* rb = smem + p.off_rows + (row & 3) * STEM_ROW_BYTES;
#pragma unroll
            for (int xx = 0; xx < 3; ++xx) {
              const int xpix = xx == 0 ? x0 : xx == 1 ? 2 * pp : x2;
              const uint4 v = *reinterpret_cast<const uint4*>(rb + xpix * 128 + ((pc ^ (xpix & 7)) << 4));
              m.x = __vmaxs2(m.x, v.x); m.y = __vmaxs2(m.y, v.y); m.z = __vmaxs2(m.z, v.z); m.w = __vmaxs2(m.w, v.w);
            }
          }
          const int y[8] = {(int)(m.x & 0xFFFF), (int)(m.x >> 16), (int)(m.y & 0xFFFF), (int)(m.y >> 16),
                            (int)(m.z & 0xFFFF), (int)(m.z >> 16), (int)(m.w & 0xFFFF), (int)(m.w >> 16)};
          if (p.y_bits == 16) {                     // [Wp pixels][128 B], SWIZZLE_128B
            *reinterpret_cast<uint4*>(smem + y_off + tile_piece_off(128, pp, pc)) = m;
          } else {                                  // int32: [2 chunks of 32 channels][Wp pixels][128 B]
            const int chunk = pc >> 2, piece = (pc & 3) * 2;
            const int yrow = chunk * p.Wp + pp;
            *reinterpret_cast<uint4*>(smem + y_off + tile_piece_off(128, yrow, piece)) = make_uint4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<uint4*>(smem + y_off + tile_piece_off(128, yrow, piece + 1)) = make_uint4(y[4], y[5], y[6], y[7]);
          }
          if (p.low_bits) {
            int q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              q[k] = clampi(__double2loint(__fma_rn(__hiloint2double(0x43300000, y[k]), low_M, low_C)), p.low_lo, p.low_hi);
            const uint32_t w0 = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
            const uint32_t w1 = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
            if (p.low_bits == 8) *reinterpret_cast<uint2*>(smem + low_off + tile_piece_off(64, pp, pc >> 1) + (pc & 1) * 8) = make_uint2(w0, w1);
            else *reinterpret_cast<uint32_t*>(smem + low_off + tile_piece_off(32, pp, pc >> 2) + (pc & 3) * 4) = pack_nibbles8(w0, w1);
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(STEM_EPI_WARPS * 32));
        if (elect_x) {
          const int row0 = (n_img * p.Hp + pr) * p.Wp;
          if (p.y_bits == 16) tma_store_2d(&maps.y, 0, row0, smem_base + y_off);
          else tma_store_3d(&maps.y, 0, row0, 0, smem_base + y_off);
          if (p.low_bits) tma_store_2d(&maps.low, 0, row0, smem_base + low_off);
          bulk_commit();
        }
      }
    }
    if (elect_x) bulk_wait_all();
    if (bad) atomicOr(p.status, HAWQ_FLAG_BAD_RATIO);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == STEM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace hawq

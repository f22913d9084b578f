// 3x3 stride-1 pad-1 convolution with the A operand read IN PLACE from a zero-padded input patch, weights stationary.
//
// Output positions are enumerated in the PADDED coordinate system of one image: a tile covers R output rows y0 .. y0 + R - 1
// and all W + 2 padded columns, position p = (y - y0) * (W + 2) + x, p < R * (W + 2) <= 128; columns x >= W are waste
// (2 / (W + 2) of the MMA rows).  One 4-D TMA box {64 channels, W + 2, R + 2, 1 image} at (c0, -1, y0 - 1, n) lands the
// input rows y0 - 1 .. y0 + R with the left / right / top / bottom padding zero-filled by the TMA unit, as a matrix of
// (R + 2) * (W + 2) rows x 64 B in the UMMA K-major SWIZZLE_64B layout.  In that matrix the im2col row of position p for
// tap (kh, kw) is row p + kh * (W + 2) + kw: a constant row shift.  So tcgen05.mma reads its A operand straight from the
// patch with the descriptor start address advanced by (kh * (W + 2) + kw) * 64 bytes (the swizzle is a function of the
// shared-memory address bits, tools/probe_b200.cu part A) - no im2col copy, no producer warps, no per-k-tile barriers:
// per (tile, 64-channel chunk) one TMA load, one barrier wait, 18 MMAs (9 taps x K = 2 x 32), one commit.
// L2 -> SM traffic is ~ (R + 2) / R of the input instead of 9x.
//
// The weights of the CTA's channel block stay in shared memory: one 3-D TMA box {64 B, BN rows, 9 taps} per 64-channel chunk,
// straight from the OHWI tensor, lands them as [chunk][tap][BN][64 B] SWIZZLE_64B blocks (a few large copies: the copy engine
// retires about one copy per 170-280 ns per SM whatever its size, tools/probe_b200.cu parts B-D).  A CTA keeps its channel
// block and walks over image row groups.
//
//   producer warp(s)   one elected lane: weight load, patch TMA loads (A4: four warps, which also expand the packed 4-bit
//                      patches to int8 once per patch instead of once per tap)
//   MMA warp           one elected lane issues tcgen05.mma kind::i8 into one of two TMEM accumulators
//   16 epilogue warps  four per TMEM lane quarter, BN / 4 columns each (the epilogue is instruction-latency bound: with two
//                      warps per scheduler 60-80 % of the cycles had no eligible warp, profiles/r02): tcgen05.ld, exact
//                      FP64-FMA dyadic requantisation (+bias, ReLU, clamp by saturating packs) -> int8 / packed uint4,
//                      every thread stores its own row (rows of waste columns are skipped)
// Semantics: QuantBnConv2d / QuantConv2d + QuantAct case 0, reference quant_modules.py:440-494, quant_utils.py:390-413.
#pragma once
#include "tc_ptx.cuh"

namespace hawq {

struct HaloParams {
  const hawq_chan* chan;
  uint8_t* out;              // NHWC, out_bits 8 (int8) or 4 (packed, hawq nibble order)
  int32_t* status;
  int N, H, W, Cout;
  int chunks;                // Cin / 64
  int R;                     // output rows per tile
  int wp;                    // W + 2
  int tiles_per_img;         // ceil(H / R)
  int m_tiles, n_tiles, ctas_per_n;
  int patch_bytes;           // bytes of one TMA box: (R + 2) * (W + 2) * (A4 ? 32 : 64)
  int patch_alloc;           // bytes per int8 patch buffer (>= (128 + 2 * wp + 2) * 64, multiple of 1024)
  int packed_alloc;          // A4: bytes per packed patch buffer (multiple of 1024)
  int npb, nkb;              // int8 patch buffers, packed patch buffers (A4)
  int relu, out_bits, lo, hi;
  long long* trace;          // debug timeline (hawq_debug_halo_trace): [4 roles][64][4] clock64 stamps of CTA 0, or null
  int w_rank3;               // weights tensor map: 1 = {Cin, Cout, 9 taps} 3-D view, 0 = plain [Cout][K] matrix
  int off_patch, off_packed, off_out, off_cst, off_bar;   // shared-memory carve-up (bytes from the 1024-aligned base; weights at 0)
};

constexpr int HALO_EPI_WARPS = 16;
__host__ __device__ constexpr int halo_producer_warps(bool a4) { return a4 ? 4 : 1; }
__host__ __device__ constexpr int halo_threads(bool a4) { return (halo_producer_warps(a4) + 1 + HALO_EPI_WARPS) * 32; }   // 576 / 672
constexpr int HALO_MAX_BUFS = 4;



template <int BN, bool A4, bool TRACE = false>
__global__ void __launch_bounds__(halo_threads(A4), 1) conv_halo_kernel(const HaloParams p, const __grid_constant__ CUtensorMap xmap,
                                                                        const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap omap) {
  constexpr int B_STAGE = BN * 64;
  constexpr int NPW = halo_producer_warps(A4);   // producer (+ converter) warps
  constexpr int MMA_WARP = NPW;
  constexpr int EPI_WARP0 = NPW + 1;
  constexpr int CW = BN / 4;               // columns per epilogue warp: 16 / 32
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  double2* sCst = reinterpret_cast<double2*>(smem + p.off_cst);
  const uint32_t bar_base = smem_base + p.off_bar;
  const uint32_t b_full = bar_base;
  auto pfull = [&](int b) { return bar_base + 8u * (1 + b); };
  auto pempty = [&](int b) { return bar_base + 8u * (1 + HALO_MAX_BUFS + b); };
  auto kfull = [&](int b) { return bar_base + 8u * (1 + 2 * HALO_MAX_BUFS + b); };
  auto kempty = [&](int b) { return bar_base + 8u * (1 + 3 * HALO_MAX_BUFS + b); };
  auto tfull = [&](int b) { return bar_base + 8u * (1 + 4 * HALO_MAX_BUFS + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (3 + 4 * HALO_MAX_BUFS + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (5 + 4 * HALO_MAX_BUFS));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto stamp = [&](int role, uint32_t idx, int ev) {
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && idx < 64) p.trace[(role * 64 + idx) * 4 + ev] = clock64();
    }
  };
  // tiles of this CTA: mt = slot, slot + ctas_per_n, ...; (image, row group) advance incrementally (no division per tile)
  const int step_n = p.ctas_per_n / p.tiles_per_img, step_t = p.ctas_per_n - step_n * p.tiles_per_img;
  const int my_tiles = (blockIdx.x / p.n_tiles < p.m_tiles) ? (p.m_tiles - 1 - blockIdx.x / p.n_tiles) / p.ctas_per_n + 1 : 0;
  const int nt = blockIdx.x % p.n_tiles, slot = blockIdx.x / p.n_tiles;
  const int n0 = nt * BN;
  const int KT = 9 * p.chunks;

  if (tid == 0) {
    mbar_init(b_full, 1);
    for (int b = 0; b < HALO_MAX_BUFS; ++b) {
      mbar_init(pfull(b), A4 ? 4 : 1);        // A4: one arrival per converter warp
      mbar_init(pempty(b), 1);
      mbar_init(kfull(b), 1);
      mbar_init(kempty(b), 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), HALO_EPI_WARPS);
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
    // weights and per-channel constants are plan-time data: fetched before waiting for the previous kernel of the stream
    if (warp == 0 && elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)KT * B_STAGE);
      if (p.w_rank3) {
        for (int c = 0; c < p.chunks; ++c)    // box {64 B of chunk c, rows n0 .. n0 + BN - 1, 9 taps} -> [c][tap][BN][64]
          tma_load_3d(smem_base + (uint32_t)c * 9u * B_STAGE, &wmap, c * 64, n0, 0, b_full);
      } else {                                // driver refused the 3-D view: one 2-D box {64 B, BN rows} per (chunk, tap)
        for (int c = 0; c < p.chunks; ++c)
          for (int t = 0; t < 9; ++t)
            tma_load_2d(smem_base + (uint32_t)(c * 9 + t) * B_STAGE, &wmap, (t * p.chunks + c) * 64, n0, b_full);
      }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if constexpr (!A4) {
      if (warp == 0 && elect_one()) {
        uint32_t b = 0, ph = 0, g = 0;
        int n_img = slot / p.tiles_per_img, ti = slot - n_img * p.tiles_per_img;
        for (int t = 0; t < my_tiles; ++t) {
          const int y0 = ti * p.R;
          for (int c = 0; c < p.chunks; ++c, ++g) {
            stamp(0, g, 0);
            mbar_wait_small(pempty(b), ph ^ 1);
            stamp(0, g, 1);
            mbar_arrive_expect_tx(pfull(b), (uint32_t)p.patch_bytes);
            tma_load_4d(smem_base + p.off_patch + b * p.patch_alloc, &xmap, c * 64, -1, y0 - 1, n_img, pfull(b));
            stamp(0, g, 2);
            if (++b == (uint32_t)p.npb) { b = 0; ph ^= 1; }
          }
          n_img += step_n; ti += step_t;
          if (ti >= p.tiles_per_img) { ti -= p.tiles_per_img; ++n_img; }
        }
      }
    } else {
      // packed 4-bit patches: TMA -> packed buffer (rows of 32 B, SWIZZLE_32B) -> expanded by these 128 threads into the int8
      // patch (rows of 64 B, SWIZZLE_64B) in the K order the permuted weights expect (per 32-channel block: low nibbles, high nibbles)
      const int total_g = my_tiles * p.chunks;
      const int rows = p.patch_bytes / 32;
      auto issue = [&](int g) {            // one elected lane of warp 0
        const int t = g / p.chunks, c = g - t * p.chunks;
        const int mt = slot + t * p.ctas_per_n;
        const int n_img = mt / p.tiles_per_img, y0 = (mt - n_img * p.tiles_per_img) * p.R;
        const int kb = g % p.nkb;
        mbar_wait_small(kempty(kb), ((g / p.nkb) & 1) ^ 1);
        mbar_arrive_expect_tx(kfull(kb), (uint32_t)p.patch_bytes);
        tma_load_4d(smem_base + p.off_packed + kb * p.packed_alloc, &xmap, c * 32, -1, y0 - 1, n_img, kfull(kb));
      };
      if (warp == 0) {
        for (int g = 0; g < p.nkb - 1 && g < total_g; ++g)
          if (elect_one()) issue(g);
        __syncwarp();
      }
      for (int g = 0; g < total_g; ++g) {
        if (warp == 0) {
          if (g + p.nkb - 1 < total_g && elect_one()) issue(g + p.nkb - 1);
          __syncwarp();
        }
        const int kb = g % p.nkb, b = g % p.npb;
        mbar_wait_small(kfull(kb), (g / p.nkb) & 1);
        mbar_wait_small(pempty(b), ((g / p.npb) & 1) ^ 1);
        const uint8_t* src = smem + p.off_packed + kb * p.packed_alloc;
        uint8_t* dst = smem + p.off_patch + b * p.patch_alloc;
        for (int r = tid; r < rows; r += 128) {
          const uint32_t p_sw = (r >> 2) & 1, a_sw = (r >> 1) & 3;
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            const uint4 wv = *reinterpret_cast<const uint4*>(src + r * 32 + ((blk ^ p_sw) << 4));
            const uint4 lo = make_uint4(wv.x & 0x0F0F0F0Fu, wv.y & 0x0F0F0F0Fu, wv.z & 0x0F0F0F0Fu, wv.w & 0x0F0F0F0Fu);
            const uint4 hi = make_uint4((wv.x >> 4) & 0x0F0F0F0Fu, (wv.y >> 4) & 0x0F0F0F0Fu, (wv.z >> 4) & 0x0F0F0F0Fu, (wv.w >> 4) & 0x0F0F0F0Fu);
            *reinterpret_cast<uint4*>(dst + r * 64 + (((2 * blk) ^ a_sw) << 4)) = lo;
            *reinterpret_cast<uint4*>(dst + r * 64 + (((2 * blk + 1) ^ a_sw) << 4)) = hi;
          }
        }
        fence_proxy_async();             // generic-proxy writes -> tcgen05.mma (async proxy) reads
        __syncwarp();
        if (lane == 0) {                 // one arrival per warp (128 arrivals on one barrier word serialise)
          mbar_arrive(pfull(b));
          mbar_arrive(kempty(kb));
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================================================================== MMA issuer
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      // This thread's instruction stream paces the kernel (18 MMAs per 64-channel chunk against ~38-64 tensor cycles each), so the
      // loop body is stripped to one add per descriptor: descriptors are handled in their own units (address >> 4, constant high
      // word), the nine tap shifts are loop-invariant registers, weight blocks advance by immediates.  A row-shifted start address
      // needs no base-offset field: the swizzle is a function of the shared-memory address bits.
      const uint32_t idesc = umma_idesc_i8(128, BN, !A4);     // packed 4-bit activations are unsigned
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw64(0) >> 32);
      constexpr uint32_t BU = B_STAGE >> 4;                    // one weight block in descriptor units
      uint32_t tap[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) tap[kh * 3 + kw] = (uint32_t)(kh * p.wp + kw) * 4u;
      const uint32_t patch0 = ((smem_base + p.off_patch) >> 4) | (1u << 16), patch_step = (uint32_t)p.patch_alloc >> 4;
      const uint32_t w0 = (smem_base >> 4) | (1u << 16);
      mbar_wait_small(b_full, 0);
      uint32_t b = 0, ph = 0, g = 0, a0 = patch0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        stamp(1, g, 0);
        mbar_wait_small(tempty(buf), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        uint32_t wl = w0;
        for (int c = 0; c < p.chunks; ++c, ++g) {
          stamp(1, g, 1);
          mbar_wait_small(pfull(b), ph);
          tc_fence_after();
          stamp(1, g, 2);
          if (c == 0) umma_i8_lohi<false>(d_tmem, a0 + tap[0], wl, desc_hi, idesc);
          else umma_i8_lohi<true>(d_tmem, a0 + tap[0], wl, desc_hi, idesc);
          umma_i8_lohi<true>(d_tmem, a0 + tap[0] + 2, wl + 2, desc_hi, idesc);
#pragma unroll
          for (int k = 1; k < 9; ++k) {
            umma_i8_lohi<true>(d_tmem, a0 + tap[k], wl + k * BU, desc_hi, idesc);
            umma_i8_lohi<true>(d_tmem, a0 + tap[k] + 2, wl + k * BU + 2, desc_hi, idesc);
          }
          umma_commit(pempty(b));
          stamp(1, g, 3);
          wl += 9 * BU;
          a0 += patch_step;
          if (++b == (uint32_t)p.npb) { b = 0; ph ^= 1; a0 = patch0; }
        }
        umma_commit(tfull(buf));
      }
    }
  } else {
    // =============================================================================== epilogue (16 warps)
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                      // column group of CW columns (4 consecutive warps cover the 4 quarters)
    constexpr double kMagic = 6755399441055744.0, kOffS = 4503601774854144.0;
    const int q_lo = p.relu ? max(p.lo, 0) : p.lo, q_hi = p.hi;
    // clamp by saturating packs where the range allows: [0, hi <= 255] -> unsigned byte saturation + per-byte min;
    // [-128, 127] -> signed byte saturation; anything else -> two integer min / max per value
    const int clamp_mode = (q_lo == 0 && q_hi >= 0 && q_hi <= 255) ? 1 : (q_lo == -128 && q_hi == 127) ? 2 : 0;
    const uint32_t hi4 = (uint32_t)(q_hi & 0xFF) * 0x01010101u;
    int bad = 0;
    // per-channel constants of this CTA's channel block (plan-time data)
    for (int i = tid - EPI_WARP0 * 32; i < BN; i += HALO_EPI_WARPS * 32) {
      const hawq_chan ch = p.chan[n0 + i];
      sCst[i] = make_double2(kOffS - (double)ch.bias, dyadic_to_double(ch.m, ch.e));
      bad |= !(ch.m == 0u || ch.e >= 31) | (ch.bias >= (1 << 29)) | (ch.bias <= -(1 << 29));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(HALO_EPI_WARPS * 32));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // this thread's output position is the same in every tile: position pos = TMEM lane, (y, x) inside the row group
    const int pos = quarter * 32 + lane;
    const double2* cst = sCst + cg * CW;
    // CW == 16: the thread's 16 channel constants live in registers (bias folded into an integer add, ratio as a double) instead
    // of one broadcast LDS.128 per value: a broadcast LDS.128 occupies the shared-memory pipe for 4 cycles per warp like any other
    // LDS.128 (tools/probe_mma.cu), 16 warps x 16 of them were a third of the tile time.  (CW == 32 has no registers for that.)
    constexpr bool REGC = (CW == 16);
    uint32_t bx[REGC ? CW : 1];
    double mm[REGC ? CW : 1];
    if constexpr (REGC) {
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const hawq_chan ch = p.chan[n0 + cg * CW + j];
        bx[j] = (uint32_t)ch.bias + 0x80000000u;
        mm[j] = dyadic_to_double(ch.m, ch.e);
      }
    }
    const bool elect_x = (ew == 0 && lane == 0);     // issues the TMA stores
    int n_img = slot / p.tiles_per_img, ti = slot - n_img * p.tiles_per_img;
    for (int t = 0; t < my_tiles; ++t) {
      const uint32_t buf = t & 1;
      if (ew == 0 && lane == 0) stamp(2, t, 0);
      mbar_wait_small(tfull(buf), (t >> 1) & 1);
      tc_fence_after();
      if (ew == 0 && lane == 0) stamp(2, t, 1);
      uint32_t acc[CW];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + cg * CW;
      if constexpr (CW == 32) tmem_ld32(taddr, acc);
      else tmem_ld16(taddr, acc);
      tmem_ld_wait();
      // the accumulator is in registers: hand the TMEM buffer back to the MMA warp before doing the arithmetic
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
      if (ew == 0 && lane == 0) stamp(2, t, 2);
      int q[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        if constexpr (REGC) {       // 2^52 + (acc + bias + 2^31) - (2^52 + 2^31): exact, |acc + bias| < 2^31 by the bias bound
          const double d = __hiloint2double(0x43300000, acc[j] + bx[j]) - kOffS;
          q[j] = __double2loint(__fma_rn(d, mm[j], kMagic));
        } else {
          const double2 cm = cst[j];
          const double d = __hiloint2double(0x43300000, acc[j] ^ 0x80000000) - cm.x;
          q[j] = __double2loint(__fma_rn(d, cm.y, kMagic));
        }
      }
      if (ew == 0 && lane == 0) stamp(3, t, 0);
      uint32_t w[CW / 4];
      if (clamp_mode == 1) {              // [0, hi]: unsigned byte saturation, then a per-byte min
#pragma unroll
        for (int j = 0; j < CW; j += 4) {
          uint32_t hi, out;
          asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q[j + 3]), "r"(q[j + 2]), "r"(0));
          asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(q[j + 1]), "r"(q[j]), "r"(hi));
          w[j / 4] = __vminu4(out, hi4);
        }
      } else if (clamp_mode == 2) {       // [-128, 127]: signed byte saturation
#pragma unroll
        for (int j = 0; j < CW; j += 4) {
          uint32_t hi;
          asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q[j + 3]), "r"(q[j + 2]), "r"(0));
          asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w[j / 4]) : "r"(q[j + 1]), "r"(q[j]), "r"(hi));
        }
      } else {
#pragma unroll
        for (int j = 0; j < CW; j += 4)
          w[j / 4] = __byte_perm(__byte_perm(clampi(q[j], q_lo, q_hi), clampi(q[j + 1], q_lo, q_hi), 0x0040),
                                 __byte_perm(clampi(q[j + 2], q_lo, q_hi), clampi(q[j + 3], q_lo, q_hi), 0x0040), 0x5410);
      }
      // stage the tile ([position][BN * bits / 8 bytes], swizzled by row length) and store it with ONE 4-D TMA box
      // {row bytes, W + 2, R, 1} at (channel block, 0, first row, image): the waste columns x >= W and rows past the image are
      // out of bounds of the output tensor and are clipped by the TMA unit
      const int rb_out = BN * p.out_bits / 8;                  // 128 / 64 / 32
      if (elect_x) bulk_wait_read_all();                        // the previous tile's store has finished reading the staging tile
      if (ew == 0 && lane == 0) stamp(3, t, 1);
      asm volatile("bar.sync 1, %0;" ::"n"(HALO_EPI_WARPS * 32));
      if (ew == 0 && lane == 0) stamp(3, t, 2);
      uint8_t* lt = smem + p.off_out;
      if (p.out_bits == 8) {
#pragma unroll
        for (int j = 0; j < CW / 16; ++j)
          *reinterpret_cast<uint4*>(lt + tile_piece_off(rb_out, pos, cg * (CW / 16) + j)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      } else {            // hawq nibble order: per 8 channels, byte j = c_j | c_{j+4} << 4
#pragma unroll
        for (int j = 0; j < CW / 16; ++j) {
          const int boff = cg * (CW / 2) + j * 8;
          *reinterpret_cast<uint2*>(lt + tile_piece_off(rb_out, pos, boff >> 4) + (boff & 8)) =
              make_uint2(pack_nibbles8(w[4 * j], w[4 * j + 1]), pack_nibbles8(w[4 * j + 2], w[4 * j + 3]));
        }
      }
      fence_proxy_async();
      if (ew == 0 && lane == 0) stamp(3, t, 3);
      asm volatile("bar.sync 1, %0;" ::"n"(HALO_EPI_WARPS * 32));
      if (elect_x) {
        tma_store_4d(&omap, n0 * p.out_bits / 8, 0, ti * p.R, n_img, smem_base + p.off_out);
        bulk_commit();
      }
      if (ew == 0 && lane == 0) stamp(2, t, 3);
      n_img += step_n; ti += step_t;
      if (ti >= p.tiles_per_img) { ti -= p.tiles_per_img; ++n_img; }
    }
    if (elect_x) bulk_wait_all();
    if (bad) atomicOr(p.status, HAWQ_FLAG_BAD_RATIO);
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

// 1x1 stride-1 convolution (a plain [M][K] x [Cout][K]^T integer GEMM) with stationary weights and few, large TMA copies.
//
// What bounds these layers on B200 is not HBM, L2 or the tensor pipe but the per-operation cost of the copy engine and of the
// single-thread instruction streams (tools/probe_b200.cu: one TMA / bulk copy retires per ~170-290 ns per SM whatever its size;
// profiles/r02/halo_trace_*.txt).  So:
//   * a CTA keeps its channel block: the whole [BN][K] weight slab is loaded once (3-D TMA boxes {64 B, BN rows, <= 8 k-tiles}
//     straight from the OHWI / [Cout][K] matrix) and the CTA walks over row tiles;
//   * activations arrive as ONE 3-D TMA box per stage: {64 B, 128 rows, KC k-tiles} -> [k-tile][128][64 B] SWIZZLE_64B blocks
//     (up to 32 KB per copy) instead of one 8 KB copy per k-tile;
//   * the uint16 residual tile of the case-1 epilogue is one box {128 B, 128 rows, BN / 64} per tile (was 8); outputs are staged
//     in shared memory (swizzled tiles) and leave with ONE TMA store per tile and output tensor (rows written by the threads
//     themselves - 16 to 64 B per row and thread - ran the store path at a fraction of its rate: profiles/r02);
//   * one elected lane issues the MMAs with descriptors advanced by constants (one add each);
//   * 16 epilogue warps (four per TMEM lane quarter), accumulator released to the MMA warp right after tcgen05.ld.
// Epilogues: REQUANT -> int8 / packed uint4 (QuantAct case 0, quant_utils.py:390-413) and RESIDUAL uint16-in / uint16-out with the
// next unit's low-bit activation (case 1, quant_utils.py:416-456 + the following quant_act), same arithmetic as conv_tc.cuh.
// A4: packed 4-bit activations ({32 B, 128 rows, KC} boxes) are expanded to int8 by four converter warps, once per stage.
#pragma once
#include "tc_ptx.cuh"

namespace hawq {

struct C1Params {
  const hawq_chan* chan;
  uint8_t* out;              // REQUANT: NHWC int8 / packed uint4.  RESIDUAL: the uint16 stream
  uint8_t* out_low;          // RESIDUAL: low-bit activation of the next unit (or null)
  int32_t* status;
  int M, Cout;
  int KT;                    // K / 64
  int KC;                    // k-tiles per activation stage (divides KT, <= 4)
  int NS;                    // activation stages
  int m_tiles, n_tiles, ctas_per_n;
  int w_boxes, w_box_kt;     // weight slab = w_boxes boxes of w_box_kt k-tiles
  int relu, out_bits, lo, hi;                                  // REQUANT
  uint32_t res_m; int res_e;                                   // RESIDUAL: scalar ratio of the uint16 stream
  int low_bits; uint32_t low_m; int low_e, low_lo, low_hi;     // RESIDUAL: low-bit copy
  int sat_pack;
  int off_a, off_packed, off_res, off_y, off_low, off_cst, off_bar;   // shared-memory carve-up (weights at 0)
  long long* trace;          // debug timeline (HAWQ_B200_HALO_TRACE=1): [4 roles][48 tiles][4] clock64 stamps of CTA 0, or null
};

constexpr int C1_EPI_WARPS = 16;
constexpr int C1_MAX_STAGES = 4;
constexpr int C1_REQ = 0, C1_RES = 1;
__host__ __device__ constexpr int c1_producer_warps(bool a4) { return a4 ? 4 : 1; }
// warps: producers (1, or 4 converters for packed input) | MMA issuer | residual loader (RESIDUAL epilogue only) | 16 epilogue warps
__host__ __device__ constexpr int c1_threads(bool a4, int epi) { return (c1_producer_warps(a4) + 1 + (epi == C1_RES ? 1 : 0) + C1_EPI_WARPS) * 32; }

struct alignas(64) C1Maps {
  CUtensorMap a;     // activations {64 | 32 B, M rows (pitch K bytes), KT}: box {64 | 32, 128, KC}
  CUtensorMap w;     // weights     {64 B, Cout rows (pitch K), KT}:        box {64, BN, w_box_kt}
  CUtensorMap res;   // uint16 stream {128 B, M rows (pitch 2 Cout), 2 Cout / 128}: box {128, 128, BN / 64}, SWIZZLE_128B
  CUtensorMap y;     // uint16 stream out, same geometry
  CUtensorMap low;   // 8 / 4-bit output [M][Cout * bits / 8]: box {BN * bits / 8, 128}, swizzle by row length
};

template <int BN, int EPI, bool WIDE, bool A4>
__global__ void __launch_bounds__(c1_threads(A4, EPI), 1) conv1x1_kernel(const C1Params p, const __grid_constant__ C1Maps maps) {
  constexpr int B_STAGE = BN * 64;           // one weight k-tile
  constexpr int A_TILE = 128 * 64;           // one activation k-tile (int8)
  constexpr int NPW = c1_producer_warps(A4);
  constexpr int MMA_WARP = NPW, RES_WARP = NPW + 1, EPI_WARP0 = NPW + 1 + (EPI == C1_RES ? 1 : 0);
  constexpr int CW = BN / 4;                 // columns per epilogue warp: 16 / 32
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int RES_BYTES = 128 * BN * 2;    // residual tile (uint16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  double2* sCst = reinterpret_cast<double2*>(smem + p.off_cst);
  const uint32_t bar_base = smem_base + p.off_bar;
  const uint32_t b_full = bar_base;
  auto afull = [&](int s) { return bar_base + 8u * (1 + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (1 + C1_MAX_STAGES + s); };
  auto kfull = [&](int s) { return bar_base + 8u * (1 + 2 * C1_MAX_STAGES + s); };    // A4: packed stage landed
  auto kempty = [&](int s) { return bar_base + 8u * (1 + 3 * C1_MAX_STAGES + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (1 + 4 * C1_MAX_STAGES + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (3 + 4 * C1_MAX_STAGES + b); };
  auto rfull = [&](int b) { return bar_base + 8u * (5 + 4 * C1_MAX_STAGES + b); };
  auto rempty = [&](int b) { return bar_base + 8u * (7 + 4 * C1_MAX_STAGES + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_bar + 8 * (9 + 4 * C1_MAX_STAGES));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto stamp = [&](int role, int idx, int ev) {      // roles: 0 producer / converter, 1 MMA, 2 epilogue, 3 residual loader
    if (p.trace != nullptr && blockIdx.x == 0 && idx < 48) p.trace[(role * 48 + idx) * 4 + ev] = clock64();
  };
  const int nt = blockIdx.x % p.n_tiles, slot = blockIdx.x / p.n_tiles;
  const int n0 = nt * BN;
  const int my_tiles = (slot < p.m_tiles) ? (p.m_tiles - 1 - slot) / p.ctas_per_n + 1 : 0;
  const int stages_per_tile = p.KT / p.KC;

  if (tid == 0) {
    mbar_init(b_full, 1);
    for (int s = 0; s < C1_MAX_STAGES; ++s) {
      mbar_init(afull(s), A4 ? 4 : 1);        // A4: one arrival per converter warp
      mbar_init(aempty(s), 1);
      mbar_init(kfull(s), 1);
      mbar_init(kempty(s), 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), C1_EPI_WARPS);
      mbar_init(rfull(b), 1);
      mbar_init(rempty(b), C1_EPI_WARPS);
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
    // weights are plan-time data: fetched before waiting for the previous kernel of the stream
    if (warp == 0 && elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)p.KT * B_STAGE);
      for (int i = 0; i < p.w_boxes; ++i)
        tma_load_3d(smem_base + (uint32_t)(i * p.w_box_kt) * B_STAGE, &maps.w, 0, n0, i * p.w_box_kt, b_full);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t stage_bytes = (uint32_t)p.KC * (A4 ? A_TILE / 2 : A_TILE);
    if constexpr (!A4) {
      if (warp == 0 && elect_one()) {
        uint32_t s = 0, ph = 0;
        for (int t = 0; t < my_tiles; ++t) {
          const int m0 = (slot + t * p.ctas_per_n) * 128;
          for (int k0 = 0; k0 < p.KT; k0 += p.KC) {
            if (k0 == 0) stamp(0, t, 0);
            mbar_wait_small(aempty(s), ph ^ 1);
            if (k0 == 0) stamp(0, t, 1);
            mbar_arrive_expect_tx(afull(s), stage_bytes);
            tma_load_3d(smem_base + p.off_a + s * (uint32_t)(p.KC * A_TILE), &maps.a, 0, m0, k0, afull(s));
            if (++s == (uint32_t)p.NS) { s = 0; ph ^= 1; }
          }
        }
      }
    } else {
      // packed 4-bit rows: TMA -> packed stage ([k-tile][128][32 B], SWIZZLE_32B) -> these 128 threads expand every row to int8 in the
      // K order the permuted weights expect (per 32-channel block: low nibbles, high nibbles) -> [k-tile][128][64 B] SWIZZLE_64B
      const int total_g = my_tiles * stages_per_tile;
      auto issue = [&](int g) {            // one elected lane of warp 0
        const int t = g / stages_per_tile, k0 = (g - t * stages_per_tile) * p.KC;
        const int m0 = (slot + t * p.ctas_per_n) * 128;
        const int ks = g % p.NS;
        mbar_wait_small(kempty(ks), ((g / p.NS) & 1) ^ 1);
        mbar_arrive_expect_tx(kfull(ks), stage_bytes);
        tma_load_3d(smem_base + p.off_packed + ks * (uint32_t)(p.KC * A_TILE / 2), &maps.a, 0, m0, k0, kfull(ks));
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
        const bool st0 = tid == 0 && g % stages_per_tile == 0;
        if (st0) stamp(0, g / stages_per_tile, 0);
        mbar_wait_small(kfull(s), (g / p.NS) & 1);
        if (st0) stamp(0, g / stages_per_tile, 1);
        mbar_wait_small(aempty(s), ((g / p.NS) & 1) ^ 1);
        if (st0) stamp(0, g / stages_per_tile, 2);
        const uint8_t* src = smem + p.off_packed + s * (p.KC * A_TILE / 2);
        uint8_t* dst = smem + p.off_a + s * (p.KC * A_TILE);
        const int r = tid;                                      // one row per thread, every k-tile of the stage
        const uint32_t p_sw = (r >> 2) & 1, a_sw = (r >> 1) & 3;
        // all loads of the stage first (KC <= 4 k-tiles x 2 vectors in flight), then the expansion: this warp is alone on its
        // scheduler, so instruction-level parallelism is what hides the shared-memory latency
        uint4 wv[4][2];
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
          if (kt < p.KC) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) wv[kt][blk] = *reinterpret_cast<const uint4*>(src + kt * (A_TILE / 2) + r * 32 + ((blk ^ p_sw) << 4));
          }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
          if (kt < p.KC) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              const uint4 v = wv[kt][blk];
              const uint4 lo = make_uint4(v.x & 0x0F0F0F0Fu, v.y & 0x0F0F0F0Fu, v.z & 0x0F0F0F0Fu, v.w & 0x0F0F0F0Fu);
              const uint4 hi = make_uint4((v.x >> 4) & 0x0F0F0F0Fu, (v.y >> 4) & 0x0F0F0F0Fu, (v.z >> 4) & 0x0F0F0F0Fu, (v.w >> 4) & 0x0F0F0F0Fu);
              *reinterpret_cast<uint4*>(dst + kt * A_TILE + r * 64 + (((2 * blk) ^ a_sw) << 4)) = lo;
              *reinterpret_cast<uint4*>(dst + kt * A_TILE + r * 64 + (((2 * blk + 1) ^ a_sw) << 4)) = hi;
            }
          }
        fence_proxy_async();             // generic-proxy writes -> tcgen05.mma (async proxy) reads
        __syncwarp();
        if (lane == 0) {                 // one arrival per warp (128 arrivals on one barrier word serialise)
          mbar_arrive(afull(s));
          mbar_arrive(kempty(s));
        }
        if (st0) stamp(0, g / stages_per_tile, 3);
      }
    }
  } else if (EPI == C1_RES && warp == RES_WARP) {
    // =============================================================================== residual loader (case-1 epilogue operand)
    // its own warp: waiting for a free buffer must stall neither the activation loads nor the converters
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      for (int t = 0; t < my_tiles; ++t) {
        const int rb = t & 1;
        stamp(3, t, 0);
        mbar_wait_small(rempty(rb), ((t >> 1) & 1) ^ 1);
        stamp(3, t, 1);
        mbar_arrive_expect_tx(rfull(rb), RES_BYTES);
        tma_load_3d(smem_base + p.off_res + rb * RES_BYTES, &maps.res, 0, (slot + t * p.ctas_per_n) * 128, n0 / 64, rfull(rb));
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================================================================== MMA issuer
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_i8(128, BN, !A4);     // packed 4-bit activations are unsigned
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw64(0) >> 32);
      constexpr uint32_t BU = B_STAGE >> 4, AU = A_TILE >> 4;  // one k-tile of weights / activations in descriptor units
      const uint32_t a_base = ((smem_base + p.off_a) >> 4) | (1u << 16), a_step = (uint32_t)p.KC * AU;
      const uint32_t w_base = (smem_base >> 4) | (1u << 16);
      mbar_wait_small(b_full, 0);
      uint32_t s = 0, ph = 0, a0 = a_base;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        stamp(1, t, 0);
        mbar_wait_small(tempty(buf), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        stamp(1, t, 1);
        const uint32_t d_tmem = tmem_base + buf * BN;
        uint32_t wl = w_base;
        for (int k0 = 0; k0 < p.KT; k0 += p.KC) {
          mbar_wait_small(afull(s), ph);
          tc_fence_after();
          if (k0 == 0) stamp(1, t, 2);
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
        umma_commit(tfull(buf));
        stamp(1, t, 3);
      }
    }
  } else {
    // =============================================================================== epilogue (16 warps)
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int cg = ew >> 2;                      // column group of CW columns
    constexpr double kMagic = 6755399441055744.0, kOffS = 4503601774854144.0, kOffU = 4503599627370496.0;
    int bad = 0, ovf = 0, ymax = 0;
    for (int i = tid - EPI_WARP0 * 32; i < BN; i += C1_EPI_WARPS * 32) {     // plan-time data
      const hawq_chan ch = p.chan[n0 + i];
      sCst[i] = make_double2(kOffS - (double)ch.bias, dyadic_to_double(ch.m, ch.e));
      bad |= !(ch.m == 0u || ch.e >= (WIDE ? 11 : 31)) | (ch.bias >= (1 << 29)) | (ch.bias <= -(1 << 29));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int row = quarter * 32 + lane;
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

    const bool elect_x = (ew == 0 && lane == 0);     // issues the TMA stores of the epilogue
    if constexpr (EPI == C1_REQ) {
      const int q_lo = p.relu ? max(p.lo, 0) : p.lo, q_hi = p.hi;
      const int clamp_mode = (q_lo == 0 && q_hi >= 0 && q_hi <= 255) ? 1 : (q_lo == -128 && q_hi == 127) ? 2 : 0;
      const uint32_t hi4 = (uint32_t)(q_hi & 0xFF) * 0x01010101u;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        mbar_wait_small(tfull(buf), (t >> 1) & 1);
        tc_fence_after();
        uint32_t acc[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + cg * CW;
        if constexpr (CW == 32) tmem_ld32(taddr, acc);
        else tmem_ld16(taddr, acc);
        tmem_ld_wait();
        tc_fence_before();               // the accumulator is in registers: hand the TMEM buffer back before the arithmetic
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        int q[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          if constexpr (REGC) {     // 2^52 + (acc + bias + 2^31) - (2^52 + 2^31): exact, |acc + bias| < 2^31 by the bias bound
            const double d = __hiloint2double(0x43300000, acc[j] + bx[j]) - kOffS;
            q[j] = __double2loint(__fma_rn(d, mm[j], kMagic));
          } else {
            const double2 cm = cst[j];
            const double d = __hiloint2double(0x43300000, acc[j] ^ 0x80000000) - cm.x;
            q[j] = __double2loint(__fma_rn(d, cm.y, kMagic));
          }
        }
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
        // stage this thread's CW * bits / 8 bytes in the swizzled output tile; one TMA store per tile (rows >= M are clipped)
        const int rb_out = BN * p.out_bits / 8;                  // tile row bytes: 128 / 64 / 32
        if (elect_x) bulk_wait_read_all();                        // the previous tile's store has finished reading the staging tile
        asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32));
        uint8_t* lt = smem + p.off_low;
        if (p.out_bits == 8) {
#pragma unroll
          for (int j = 0; j < CW / 16; ++j)
            *reinterpret_cast<uint4*>(lt + tile_piece_off(rb_out, row, cg * (CW / 16) + j)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        } else {            // hawq nibble order: per 8 channels, byte j = c_j | c_{j+4} << 4
#pragma unroll
          for (int j = 0; j < CW / 16; ++j) {
            const int boff = cg * (CW / 2) + j * 8;             // byte offset of these 16 channels in the row
            *reinterpret_cast<uint2*>(lt + tile_piece_off(rb_out, row, boff >> 4) + (boff & 8)) =
                make_uint2(pack_nibbles8(w[4 * j], w[4 * j + 1]), pack_nibbles8(w[4 * j + 2], w[4 * j + 3]));
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32));
        if (elect_x) {
          tma_store_2d(&maps.low, n0 * p.out_bits / 8, (slot + t * p.ctas_per_n) * 128, smem_base + p.off_low);
          bulk_commit();
        }
      }
      if (elect_x) bulk_wait_all();
    } else {
      // ---- case 1: y = max(RHE((acc + bias) * M_c) + RHE(res * res_M), 0) -> uint16 stream (sticky overflow flag), plus
      //      low = clamp(RHE(y * low_M)) for the next unit; unsigned operands use the folded one-FMA form (exact for e <= 51)
      const double low_M = dyadic_to_double(p.low_m, p.low_e), res_M = dyadic_to_double(p.res_m, p.res_e);
      const double low_C = kMagic - kOffU * low_M, res_C = kMagic - kOffU * res_M;
      const bool sat8 = p.sat_pack != 0 && p.low_bits == 8 && p.low_hi == 127 && p.low_lo <= 0;
      // ... and with clamp [<= 0, hi <= 255] (4-bit values in byte containers) the u8 saturation followed by a per-byte min
      const bool satu = p.sat_pack != 0 && p.low_bits == 8 && !sat8 && p.low_lo <= 0 && p.low_hi >= 0 && p.low_hi <= 255;
      const uint32_t hi4 = (uint32_t)(p.low_hi & 255) * 0x01010101u;
      const int l_lo = p.low_lo, l_hi = p.low_hi;
      auto ratio_ok = [](uint32_t m_, int e_) { return m_ == 0u || e_ >= (WIDE ? 11 : 31); };
      bad |= !p.relu | !ratio_ok(p.res_m, p.res_e) | (p.res_m != 0u && p.res_e > 51);
      if constexpr (WIDE) bad |= !(res_M < 16384.0);
      if (p.low_bits) bad |= !dyadic_is_fast(p.low_m, p.low_e) | (p.low_m != 0u && p.low_e > 51);
      // residual tile in shared memory: [chunk of 64 columns][128 rows][128 B], SWIZZLE_128B; this thread's CW columns
      const uint32_t r_chunk = (uint32_t)(cg * CW) / 64, r_piece0 = ((uint32_t)(cg * CW) % 64) / 8;
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        if (elect_x) stamp(2, t, 0);
        mbar_wait_small(rfull(buf), (t >> 1) & 1);
        const uint8_t* rrow = smem + p.off_res + buf * RES_BYTES + r_chunk * (128 * 128) + row * 128;
        uint4 rv[CW / 8];
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) rv[i] = *reinterpret_cast<const uint4*>(rrow + (((r_piece0 + i) ^ (uint32_t)(row & 7)) << 4));
        // (the buffer is handed back at the end of the tile, after the loaded values were consumed: an arrive issued right behind
        // the loads can overtake them, and the next tile's TMA write then races the reads - seen on hardware)
        mbar_wait_small(tfull(buf), (t >> 1) & 1);
        tc_fence_after();
        uint32_t acc[CW];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + cg * CW;
        if constexpr (CW == 32) tmem_ld32(taddr, acc);
        else tmem_ld16(taddr, acc);
        tmem_ld_wait();
        if (elect_x) stamp(2, t, 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        uint4 yo[CW / 8];
        uint32_t lw[CW / 4];
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) {           // groups of 8 channels: one 16-byte vector of residuals, one of outputs
          const uint32_t rr[4] = {rv[i].x, rv[i].y, rv[i].z, rv[i].w};
          int y[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            double qv;
            if constexpr (REGC) {
              qv = __fma_rn(__hiloint2double(0x43300000, acc[i * 8 + k] + bx[i * 8 + k]) - kOffS, mm[i * 8 + k], kMagic);
            } else {
              const double2 cm = cst[i * 8 + k];
              qv = __fma_rn(__hiloint2double(0x43300000, acc[i * 8 + k] ^ 0x80000000) - cm.x, cm.y, kMagic);
            }
            int v = __double2loint(qv);
            const int r16 = (k & 1) ? (int)(rr[k >> 1] >> 16) : (int)(rr[k >> 1] & 0xFFFF);
            const double qr = __fma_rn(__hiloint2double(0x43300000, r16), res_M, res_C);
            const int vr = __double2loint(qr);
            if constexpr (WIDE) {
              // the main term may leave int32 (ratio > 1): exact per-value check.  The identity term cannot (0 <= r < 2^16 and
              // res ratio < 2^14, checked once above: vr < 2^30), and with v capped at 2^30 the sum cannot wrap: a capped value
              // is far above 65535 and raises HAWQ_FLAG_RESIDUAL_OVERFLOW through ymax like any other overflow of the stream
              ovf |= (__double2hiint(qv) + (int)((uint32_t)v >> 31)) ^ 0x43380000;
              v = min(v, 1 << 30);
            }
            const int sum = v + vr;
            y[k] = max(sum, 0);
            ymax = max(ymax, y[k]);
          }
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[i].x) : "r"(y[1]), "r"(y[0]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[i].y) : "r"(y[3]), "r"(y[2]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[i].z) : "r"(y[5]), "r"(y[4]));
          asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(yo[i].w) : "r"(y[7]), "r"(y[6]));
          if (p.low_bits) {
            int q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = __double2loint(__fma_rn(__hiloint2double(0x43300000, y[k]), low_M, low_C));   // y >= 0
            if (sat8) {
              uint32_t h0, h1;
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(h0) : "r"(q[3]), "r"(q[2]), "r"(0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lw[2 * i]) : "r"(q[1]), "r"(q[0]), "r"(h0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(h1) : "r"(q[7]), "r"(q[6]), "r"(0));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lw[2 * i + 1]) : "r"(q[5]), "r"(q[4]), "r"(h1));
            } else if (satu) {
              uint32_t h0, h1, o0, o1;
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(h0) : "r"(q[3]), "r"(q[2]), "r"(0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(o0) : "r"(q[1]), "r"(q[0]), "r"(h0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(h1) : "r"(q[7]), "r"(q[6]), "r"(0));
              asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(o1) : "r"(q[5]), "r"(q[4]), "r"(h1));
              lw[2 * i] = __vminu4(o0, hi4);
              lw[2 * i + 1] = __vminu4(o1, hi4);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) q[k] = clampi(q[k], l_lo, l_hi);
              lw[2 * i] = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
              lw[2 * i + 1] = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
            }
          }
        }
        // the residual values were consumed: hand the buffer back (an arrive issued right behind the loads can overtake them)
        __syncwarp();
        if (lane == 0) mbar_arrive(rempty(buf));
        // stage y ([64-column chunk][128 rows][128 B], SWIZZLE_128B) and the low-bit tile; one TMA store each per tile
        if (elect_x) stamp(2, t, 2);
        if (elect_x) bulk_wait_read_all();                        // the previous tile's stores have finished reading the staging tiles
        asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32));
        uint8_t* yt = smem + p.off_y + r_chunk * (128 * 128);
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) *reinterpret_cast<uint4*>(yt + tile_piece_off(128, row, (int)r_piece0 + i)) = yo[i];
        if (p.low_bits) {
          uint8_t* lt = smem + p.off_low;
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
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32));
        if (elect_x) {
          const int m0 = (slot + t * p.ctas_per_n) * 128;
          tma_store_3d(&maps.y, 0, m0, n0 / 64, smem_base + p.off_y);
          if (p.low_bits) tma_store_2d(&maps.low, n0 * p.low_bits / 8, m0, smem_base + p.off_low);
          bulk_commit();
          stamp(2, t, 3);
        }
      }
      if (elect_x) bulk_wait_all();
      if (ymax > 65535) atomicOr(p.status, HAWQ_FLAG_RESIDUAL_OVERFLOW);
    }
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

// Persistent warp-specialised implicit-GEMM convolution on 5th-gen tensor cores (tcgen05.mma kind::i8, TMEM
// accumulators) with the HAWQ epilogues fused.  One CTA per SM loops over 128 x BN output tiles (m fastest).
//
//   warps 0-3   producers: fill a STAGES-deep shared-memory ring of A (128 x 64 B) and B (BN x 64 B) k-tiles in the UMMA
//               canonical K-major SWIZZLE_64B layout (16-byte chunk index XOR (row >> 1) & 3).
//                 B: one linear cp.async.bulk per k-tile from the re-tiled weight copy (or a TMA box of plain OHWI weights);
//                 A: TMA boxes (1x1 stride-1 int8 layers) | 3x3 stride-1 "patch mode": one TMA box per (tile, Cin chunk)
//                    brings the pixel range of all nine taps, the taps are copied shared -> shared | coalesced cp.async
//                    gather with hardware-signalled mbarrier arrival (strided layers, dual mode);
//                 packed 4-bit activations are expanded to int8 here (A4);
//   warp  4     allocates TMEM; one lane issues tcgen05.mma (M = 128, N = BN, K = 32; two per k-tile) into one of two TMEM
//               accumulator buffers (dual mode: two accumulators per buffer), tcgen05.commit releases smem stages / publishes
//               the accumulator;
//   warps 5-12  epilogue: tcgen05.ld the accumulator (thread = output row, 32 consecutive channels per load), exact FP64-FMA
//               dyadic requantisation, residual add, ReLU, low-bit copy.  REQUANT / int32 variants stage outputs in a private
//               padded slice and copy out coalesced; the uint16-stream variants (RES22, DUAL) read the residual tile and write
//               the new stream + low-bit tile as swizzled TMA boxes.
//   While the epilogue of tile i runs, the producers and the MMA warp are already working on tile i+1.
//
// Supported: a_bits 8 (signed) / 4 (unsigned, packed), epilogues REQUANT -> 4/8 bit, RESIDUAL (relu), RAW_I32, DUAL; dyadic
// ratios <= 1, or <= 2^20 with the WIDE overflow checks (promised by the caller through HAWQ_EP_RATIOS_* and re-checked here:
// a violation raises HAWQ_FLAG_BAD_RATIO / HAWQ_FLAG_REQUANT_OVERFLOW in the status word instead of producing wrong numbers).
// Everything else uses conv_igemm.cuh.  Every mbarrier wait is bounded: a protocol bug traps (launch failure) instead of
// hanging the GPU.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_igemm.cuh"
#include "tc_ptx.cuh"

namespace hawq {

// TMA tensor maps of one launch (passed as a __grid_constant__ kernel parameter)
struct alignas(64) TcMaps {
  CUtensorMap b;     // weights  [Cout][K] int8, box 64 x BN, SWIZZLE_64B
  CUtensorMap a;     // activations as a matrix [M][Cin] int8 (1x1 stride-1 layers), box 64 x 128, SWIZZLE_64B
  CUtensorMap res;   // uint16 residual stream [M][Cout], box (CW * 2 bytes) x 32 rows, SWIZZLE_128B
  CUtensorMap y;     // uint16 residual stream out, same shape
  CUtensorMap low;   // low-bit activation out [M][Cout * bits / 8], box (CW * bits / 8) x 32 rows
  CUtensorMap patch; // input pixels as a matrix [N*H*W][Cin * bits / 8], box (64 | 32 bytes) x (128 + 2W + 2) rows (3x3 patch mode)
};

constexpr int TC_BM = 128;
constexpr int TC_PRODUCER_WARPS = 4;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_MMA_WARP = TC_PRODUCER_WARPS;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1 + TC_EPI_WARPS) * 32;   // 416

// epilogue variants (compile-time): element sizes of the residual operand read into / the output staged in the slice
constexpr int TC_EPI_REQ = 0;      // REQUANT -> 4/8 bit
constexpr int TC_EPI_RAW = 1;      // RAW_I32
constexpr int TC_EPI_RES22 = 2;    // RESIDUAL: uint16 stream in, uint16 stream out
constexpr int TC_EPI_RES44 = 3;    // RESIDUAL: int32 in (stream or identity-conv accumulator), int32 out
constexpr int TC_EPI_RES42 = 4;    // RESIDUAL: int32 in, uint16 out
constexpr int TC_EPI_DUAL = 5;     // RESIDUAL whose identity operand is a second in-kernel 1x1 convolution (two TMEM accumulators), uint16 out

template <int BN, int EPI, bool A4 = false>
struct TcSmem {
  static constexpr int EW = TC_EPI_WARPS;
  // pipeline depth: RESIDUAL epilogues are epilogue-bound and need shared memory for their tiles; the others are
  // load-latency-bound and get a deep ring.  The producer keeps LAG + 1 k-tiles in flight per thread.
  static constexpr int STAGES = (EPI == TC_EPI_RES22) ? (A4 && BN == 128 ? 4 : 5) : (EPI == TC_EPI_DUAL) ? 6 : (EPI >= TC_EPI_RES44) ? 5 : (EPI == TC_EPI_RAW) ? 7 : (BN == 128 ? 8 : 10);
  static constexpr int LAG = STAGES - 2;
  static constexpr int A_STAGE = TC_BM * 64;
  static constexpr int B_STAGE = BN * 64;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int RING = STAGES * STAGE;
  static constexpr int CW = BN / (EW / 4);                                 // columns per epilogue warp (4 warps cover the 128 TMEM lanes)
  static constexpr int RES_ES = (EPI == TC_EPI_RES22) ? 2 : ((EPI == TC_EPI_RES44 || EPI == TC_EPI_RES42) ? 4 : 0);
  static constexpr int Y_ES = (EPI == TC_EPI_RES22 || EPI == TC_EPI_RES42 || EPI == TC_EPI_DUAL) ? 2 : ((EPI == TC_EPI_RES44 || EPI == TC_EPI_RAW) ? 4 : 0);
  static constexpr int SLICE_ES = RES_ES > Y_ES ? RES_ES : Y_ES;
  static constexpr bool TMA_IO = (EPI == TC_EPI_RES22 || EPI == TC_EPI_DUAL);                    // residual tile in / outputs out as swizzled TMA boxes
  static constexpr int SLICE_PITCH = TMA_IO ? CW * SLICE_ES : CW * SLICE_ES + 16;   // padded: 16-byte row-per-lane accesses conflict-free
  static constexpr int SLICE = SLICE_ES ? (TMA_IO ? 4096 : 32 * SLICE_PITCH) : 0;
  static constexpr int SLICE_BUFS = (EPI == TC_EPI_RES22) ? 3 : 1;         // RES22: 2 prefetch + 1 output; RES4x: in place; RAW: output
  static constexpr int LOW_PITCH = TMA_IO ? CW : CW + 16;
  static constexpr int LOW_SLICE = (EPI == TC_EPI_RAW) ? 0 : (TMA_IO ? 2048 : 32 * LOW_PITCH);   // TMA boxes keep 1024 B alignment
  static constexpr int SLICES_OFF = RING;
  static constexpr int LOW_OFF = SLICES_OFF + EW * SLICE * SLICE_BUFS;
  static constexpr int CST_OFF = LOW_OFF + EW * LOW_SLICE;       // double2 {Cb, M}[BN]
  static constexpr int M1_OFF = CST_OFF + BN * 16;                         // double M1[BN]  (DUAL: double2 {Cb2, M1}[BN])
  static constexpr int STG_SLOT = TC_BM * 32;                              // packed 4-bit rows of one k-tile (A4 only)
  static constexpr int STG_OFF = M1_OFF + BN * (EPI == TC_EPI_DUAL ? 16 : 8);
  static constexpr int PATCH_BUF = (EPI == TC_EPI_REQ) ? 256 * 64 : 0;      // 3x3 patch mode: two buffers of <= 256 pixel rows
  static constexpr int PATCH_OFF = STG_OFF + (A4 ? (LAG + 1) * STG_SLOT : 0);
  static constexpr int BAR_OFF = PATCH_OFF + 2 * PATCH_BUF;               // mbarriers + tmem base
  static constexpr int TOTAL = BAR_OFF + 1024 + 1024;                       // + slack for 1024 B alignment of the ring
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// ---------------------------------------------------------------------------------------------- kernel
// WIDE: dyadic ratios up to 2^20 are allowed (e >= 11).  The FMA result is then checked to be a valid int32
// (high word + sign bit of the low word must equal the high word of 1.5 * 2^52); a violation raises
// HAWQ_FLAG_REQUANT_OVERFLOW so the host can re-run on the saturating generic kernels.
// A4: activations are packed unsigned nibbles (hawq order).  They are fetched packed (half the bytes), parked in a small
// staging ring and expanded to int8 by the producer thread that owns the row, directly into the swizzled A tile in the
// K order the (host-permuted) weights expect: per 32-channel block {c0-3, c8-11, c16-19, c24-27 | c4-7, c12-15, ...}.
// Preconditions (promised via HAWQ_EP_RATIOS_*, re-checked -> HAWQ_FLAG_BAD_RATIO): ratios within the bound,
// |bias| < 2^29 (sums of two requantised terms then cannot wrap), RESIDUAL launches have relu = 1.
// Single-lane roles (TMA / bulk-copy issue, MMA issue) pick their lane with elect.sync from the converged warp: nvcc then emits
// straight-line UTMALDG / UBLKCP / UTCIMMA / UTCBAR, whereas issue from a `lane == 0` branch is wrapped in an ELECT + BRA.U.ANY
// loop per instruction.  Barrier waits use a non-unrolled bounded spin loop (~8 SASS instructions per call site).  (Round-2
// A/B on B200, ResNet-50 W8A8 batch 128: 2.539 -> 2.433 ms/step, profiles/r02/a0_variants.txt.)
template <int BN, int EPI, bool WIDE, bool A4>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvParams p, const __grid_constant__ TcMaps maps) {
  using S = TcSmem<BN, EPI, A4>;
  constexpr int EW = TC_EPI_WARPS;
  constexpr int EPI_WARP0 = TC_MMA_WARP + 1;     // first epilogue warp
  constexpr int BM = TC_BM, STAGES = S::STAGES, LAG = S::LAG;
  constexpr int CW = S::CW;                      // columns handled by one epilogue warp: 32 or 64
  constexpr int TMEM_COLS = (EPI == TC_EPI_DUAL ? 4 : 2) * BN;   // two accumulator buffers (x2 accumulators in dual mode): 128 / 256 / 512
  constexpr int ACC_STRIDE = (EPI == TC_EPI_DUAL ? 2 : 1) * BN;  // TMEM columns per buffer
  constexpr int RES_ES = S::RES_ES, Y_ES = S::Y_ES, PITCH = S::SLICE_PITCH;
  constexpr bool IS_RES = EPI >= TC_EPI_RES22;
  constexpr bool DUAL = EPI == TC_EPI_DUAL;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps shared-space provenance
  const uint32_t smem_base = smem_u32(smem);
  double2* sCst = reinterpret_cast<double2*>(smem + S::CST_OFF);
  double* sM1 = reinterpret_cast<double*>(smem + S::M1_OFF);
  double2* sCst2 = reinterpret_cast<double2*>(smem + S::M1_OFF);   // dual mode: {Cb2, M1} of the identity convolution
  const uint32_t bar_base = smem_base + S::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::BAR_OFF + 8 * (2 * STAGES + 4));
  auto res_bar = [&](int ew_, int b) { return bar_base + 8u * (2 * STAGES + 6 + ew_ * 2 + b); };   // TMA residual tiles (per epilogue warp)
  auto patch_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 6 + 2 * EW + b); };            // TMA input patches (3x3 patch mode)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto kwait = [&](uint32_t bar, uint32_t parity) { mbar_wait_small(bar, parity); };
  // the thread that issues TMA / bulk copies on behalf of the producers: a lane of warp 0 picked with elect.sync
  auto tma_issuer = [&]() -> bool { return warp == 0 && elect_one() != 0; };
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.Cout / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int KT1 = p.KH * p.KW * p.cin_chunks;                 // k-tiles of the main convolution
  const int KT = KT1 + (DUAL ? p.cin_chunks2 : 0);            // + the identity-branch convolution (dual mode)

  // ---- one-time setup ----
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), (p.tma_a && !(EPI == TC_EPI_REQ && p.patch_rows != 0)) ? 1 : TC_PRODUCER_WARPS * 32 + 1);   // + the expect_tx arrival of the TMA-issuing thread
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), EW);
      for (int w = 0; w < EW; ++w) mbar_init(res_bar(w, b), 1);
      mbar_init(patch_bar(b), 1);
    }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch: everything above touched only shared memory / TMEM / kernel parameters, so (when launched
  // with the programmatic-serialization attribute) it overlaps the tail of the previous kernel in the stream.  From here on
  // global memory is read and written: wait until the preceding grid has completed and flushed.  Both are no-ops otherwise.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < TC_PRODUCER_WARPS) {
    // =============================================================================== producers (128 threads)
    // Coalesced gather: consecutive lanes cover one row's bytes (4 lanes x 16 B = 64 B int8 row, 2 lanes x 16 B = packed
    // 4-bit row), so a warp-level cp.async touches 8 (16) cache lines instead of 32.  Each thread serves A_PASSES rows.
    constexpr int A_LPR = A4 ? 2 : 4;                       // lanes (= 16-byte chunks) per A row in global memory
    constexpr int A_PASSES = A_LPR;                         // rows per thread: 128 rows * A_LPR chunks / 128 threads
    const int a_ch = tid % A_LPR;
    const bool tma_a = !A4 && p.tma_a != 0;    // activations by TMA: only thread 0 works in this role
    uint32_t it = 0;                           // global k-tile counter (ring position)
    uint32_t pending = 0;                      // k-tiles issued but not yet signalled
    // A4: expand the 16 packed bytes (one 32-channel block) this thread loaded for each of its rows of k-tile j
    auto expand_rows = [&](uint32_t j) {
#pragma unroll
      for (int i = 0; i < A_PASSES; ++i) {
        const int row = tid / A_LPR + i * (128 / A_LPR);
        const uint32_t sw = (row >> 1) & 3;
        const uint4 w = *reinterpret_cast<const uint4*>(smem + S::STG_OFF + (j % (LAG + 1)) * S::STG_SLOT + row * 32 + a_ch * 16);
        uint8_t* dst = smem + (j % STAGES) * S::STAGE + row * 64;
        const uint4 lo = make_uint4(w.x & 0x0F0F0F0Fu, w.y & 0x0F0F0F0Fu, w.z & 0x0F0F0F0Fu, w.w & 0x0F0F0F0Fu);
        const uint4 hi = make_uint4((w.x >> 4) & 0x0F0F0F0Fu, (w.y >> 4) & 0x0F0F0F0Fu, (w.z >> 4) & 0x0F0F0F0Fu, (w.w >> 4) & 0x0F0F0F0Fu);
        *reinterpret_cast<uint4*>(dst + (((2 * a_ch) ^ sw) << 4)) = lo;
        *reinterpret_cast<uint4*>(dst + (((2 * a_ch + 1) ^ sw) << 4)) = hi;
      }
    };
    if (EPI == TC_EPI_REQ && p.patch_rows != 0) {   // (never in dual mode)
      // ---- 3x3 stride-1 pad-1: im2col from shared memory.  One TMA box per (tile, Cin chunk) brings the contiguous pixel
      // range [P0 - W - 1, P0 + 127 + W + 1] (zero-filled outside the tensor) into a patch buffer; for tap (kh, kw) output row
      // r needs patch row r + kh*W + kw.  Each producer thread copies (and for packed 4-bit input expands) its own row into
      // the swizzled A tile, writing zeros where the tap falls into the padding.  L2 -> SM traffic: ~1.3x instead of 9x.
      constexpr int PROWB = A4 ? 32 : 64;                  // bytes per patch row
      const int row = tid;
      const uint32_t a_sw = (row >> 1) & 3;
      const int chunks = p.cin_chunks;
      const uint32_t patch_bytes = (uint32_t)p.patch_rows * PROWB;
      const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const long long total_g = (long long)my_tiles * chunks;           // (tile, chunk) pairs of this CTA
      auto issue_patch = [&](long long g) {                              // thread 0 only
        const int tile = blockIdx.x + (int)(g / chunks) * gridDim.x;
        const int c = (int)(g % chunks);
        const int m0 = (tile % m_tiles) * BM;
        mbar_arrive_expect_tx(patch_bar((int)(g & 1)), patch_bytes);
        tma_load_2d(smem_base + S::PATCH_OFF + (uint32_t)(g & 1) * S::PATCH_BUF, &maps.patch, c * PROWB, m0 - p.W - 1, patch_bar((int)(g & 1)));
      };
      if (total_g > 0 && tma_issuer()) issue_patch(0);
      long long g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;
        const int m = m0 + row;
        const bool row_ok = m < p.M;
        const int mm = row_ok ? m : 0;
        const int rr = mm % (p.H * p.W);
        const int h = rr / p.W, w = rr - h * p.W;
        for (int c = 0; c < chunks; ++c, ++g) {
          if (g + 1 < total_g && tma_issuer()) issue_patch(g + 1);        // prefetch into the buffer freed one chunk ago
          kwait(patch_bar((int)(g & 1)), (uint32_t)((g >> 1) & 1));
          const uint8_t* patch = smem + S::PATCH_OFF + (g & 1) * S::PATCH_BUF;
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh, it += 3) {
            // one kernel row (kw = 0, 1, 2) per iteration: three k-tiles share the barrier waits' latency, one proxy fence
            const bool vh = row_ok && (unsigned)(h + kh - 1) < (unsigned)p.H;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int stage = (it + kw) % STAGES;
              kwait(empty_bar(stage), (((it + kw) / STAGES) & 1) ^ 1);
              if (tma_issuer()) {                                        // weights of this k-tile: K index = tap * Cin + c * 64
                const int ktile = (kh * 3 + kw) * chunks + c;
                mbar_arrive_expect_tx(full_bar(stage), S::B_STAGE);
                if (p.w_tiled) bulk_load_1d(smem_base + stage * S::STAGE + S::A_STAGE, p.w_tiled + ((size_t)(n0 / BN) * KT1 + ktile) * S::B_STAGE, S::B_STAGE, full_bar(stage));
                else tma_load_2d(smem_base + stage * S::STAGE + S::A_STAGE, &maps.b, ktile * 64, n0, full_bar(stage));
              }
            }
            constexpr int NV = A4 ? 2 : 4;                               // 16-byte vectors per patch row
            uint4 val[3][NV];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const bool v = vh && (unsigned)(w + kw - 1) < (unsigned)p.W;
              const int pr = row + kh * p.W + kw;
              const uint32_t p_sw = A4 ? ((pr >> 2) & 1) : ((pr >> 1) & 3);
#pragma unroll
              for (int ch = 0; ch < NV; ++ch) {
                val[kw][ch] = make_uint4(0, 0, 0, 0);
                if (v) val[kw][ch] = *reinterpret_cast<const uint4*>(patch + pr * PROWB + ((ch ^ p_sw) << 4));
              }
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              uint8_t* dst = smem + ((it + kw) % STAGES) * S::STAGE + row * 64;
              if constexpr (!A4) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + ((ch ^ a_sw) << 4)) = val[kw][ch];
              } else {
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                  const uint4 wv = val[kw][blk];
                  const uint4 lo = make_uint4(wv.x & 0x0F0F0F0Fu, wv.y & 0x0F0F0F0Fu, wv.z & 0x0F0F0F0Fu, wv.w & 0x0F0F0F0Fu);
                  const uint4 hi = make_uint4((wv.x >> 4) & 0x0F0F0F0Fu, (wv.y >> 4) & 0x0F0F0F0Fu, (wv.z >> 4) & 0x0F0F0F0Fu, (wv.w >> 4) & 0x0F0F0F0Fu);
                  *reinterpret_cast<uint4*>(dst + (((2 * blk) ^ a_sw) << 4)) = lo;
                  *reinterpret_cast<uint4*>(dst + (((2 * blk + 1) ^ a_sw) << 4)) = hi;
                }
              }
            }
            fence_proxy_async();
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) mbar_arrive(full_bar((it + kw) % STAGES));
          }
          asm volatile("bar.sync 2, %0;" ::"n"(TC_PRODUCER_WARPS * 32));   // every thread is done reading this patch buffer
        }
      }
    } else if (tma_a) {
        // TMA producer: one lane of warp 0 chosen with elect.sync (straight-line UTMALDG / UBLKCP instead of an ELECT loop per
        // issue), ring position tracked incrementally, the weight pointer of the channel block hoisted out of the k loop
        if (warp == 0 && elect_one()) {
          uint32_t stage = 0, phase = 0;
          for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;
            const int8_t* wsrc = p.w_tiled ? p.w_tiled + (size_t)(n0 / BN) * KT1 * S::B_STAGE : nullptr;
            for (int kt = 0; kt < KT; ++kt) {
              kwait(empty_bar(stage), phase ^ 1);
              const uint32_t a_base = smem_base + stage * S::STAGE;
              mbar_arrive_expect_tx(full_bar(stage), S::A_STAGE + S::B_STAGE);
              tma_load_2d(a_base, &maps.a, kt * 64, m0, full_bar(stage));
              if (wsrc) bulk_load_1d(a_base + S::A_STAGE, wsrc + (size_t)kt * S::B_STAGE, S::B_STAGE, full_bar(stage));
              else tma_load_2d(a_base + S::A_STAGE, &maps.b, kt * 64, n0, full_bar(stage));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
    } else {
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;   // m fastest: weights / constants change rarely
      int hi0[A_PASSES], wi0[A_PASSES], pix[A_PASSES];
      int pix2[A_PASSES];                       // dual mode: identity-conv input pixel (1x1, pad 0, stride2)
      bool a_ok[A_PASSES];
#pragma unroll
      for (int i = 0; i < A_PASSES; ++i) {
        const int m = m0 + tid / A_LPR + i * (128 / A_LPR);
        a_ok[i] = m < p.M;
        const int mm = a_ok[i] ? m : 0;
        const int n = mm / (p.Ho * p.Wo);
        const int r = mm - n * (p.Ho * p.Wo);
        const int ho = r / p.Wo, wo = r - ho * p.Wo;
        hi0[i] = ho * p.stride - p.pad;
        wi0[i] = wo * p.stride - p.pad;
        pix[i] = n * p.H * p.W;
        pix2[i] = DUAL ? (n * p.H2 + ho * p.stride2) * p.W2 + wo * p.stride2 : 0;
      }
      int c = 0, kw = 0, kh = 0;
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const int stage = it % STAGES;
        kwait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
        const uint32_t a_base = smem_base + stage * S::STAGE;
        const uint32_t b_base = a_base + S::A_STAGE;
#pragma unroll
        for (int i = 0; i < A_PASSES; ++i) {
          const int row = tid / A_LPR + i * (128 / A_LPR);
          const int hi = hi0[i] + kh, wi = wi0[i] + kw;
          bool v = a_ok[i] && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W;
          const uint8_t* src = v ? p.x + (size_t)(pix[i] + hi * p.W + wi) * p.x_pix_bytes + c * (A4 ? 32 : 64) + a_ch * 16 : p.x;
          if constexpr (DUAL) {
            if (kt >= KT1) {                   // identity branch: k-tile (kt - KT1) of the second input tensor
              v = a_ok[i];
              src = v ? p.x2 + (size_t)pix2[i] * p.x2_pix_bytes + (kt - KT1) * (A4 ? 32 : 64) + a_ch * 16 : p.x;
            }
          }
          if constexpr (!A4) cp_async_16(a_base + swz<64>(row, a_ch), src, v ? 16 : 0);
          else cp_async_16(smem_base + S::STG_OFF + (it % (LAG + 1)) * S::STG_SLOT + row * 32 + a_ch * 16, src, v ? 16 : 0);
        }
        if (tma_issuer()) {                     // weights: one TMA box (64 x BN, SWIZZLE_64B) per k-tile
          mbar_arrive_expect_tx(full_bar(stage), S::B_STAGE);
          if (DUAL && kt >= KT1) bulk_load_1d(b_base, p.w2_tiled + ((size_t)(n0 / BN) * p.cin_chunks2 + (kt - KT1)) * S::B_STAGE, S::B_STAGE, full_bar(stage));
          else if (p.w_tiled) bulk_load_1d(b_base, p.w_tiled + ((size_t)(n0 / BN) * KT1 + kt) * S::B_STAGE, S::B_STAGE, full_bar(stage));
          else tma_load_2d(b_base, &maps.b, kt * 64, n0, full_bar(stage));
        }
        if constexpr (!A4) {
          // int8 rows need no post-processing: the barrier is signalled by the copy hardware itself, the producer never
          // waits for data (the MMA thread issues the generic->async proxy fence after its barrier wait)
          cp_async_mbar_arrive_noinc(full_bar(stage));
        } else {
          cp_async_commit();
          ++pending;
          if (pending > LAG) {                  // the group issued LAG iterations ago has landed: expand it
            cp_async_wait<LAG>();
            expand_rows(it - LAG);
            fence_proxy_async();
            mbar_arrive(full_bar((it - LAG) % STAGES));
            --pending;
          }
        }
        if (++c == p.cin_chunks) { c = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
      }
    }
    // drain
    cp_async_wait<0>();
    if constexpr (A4)
      for (uint32_t j = pending; j > 0; --j) expand_rows(it - j);
    fence_proxy_async();
    for (uint32_t j = pending; j > 0; --j) mbar_arrive(full_bar((it - j) % STAGES));
    }
  } else if (warp == TC_MMA_WARP) {
    // =============================================================================== MMA issuer (one lane)
    // The issuing lane is chosen with elect.sync by the converged warp.  The ring position is tracked incrementally, the
    // shared-memory descriptors advance by constants (low word = address >> 4 | 1 << 16, the high word never changes) and the
    // generic->async proxy fence is issued only when the A tile was written by plain cp.async (the other producers fence before
    // they arrive, TMA needs none): ~15 SASS instructions per k-tile.
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_i8(BM, BN, !A4);   // packed 4-bit activations are unsigned
      uint32_t tile_iter = 0;
      {
        const bool consumer_fence = !A4 && p.tma_a == 0 && !(EPI == TC_EPI_REQ && p.patch_rows != 0);
        const uint64_t desc_hi = umma_desc_sw64(0) & 0xFFFFFFFF00000000ull;
        const uint32_t a_lo0 = (smem_base >> 4) | (1u << 16);
        uint32_t stage = 0, phase = 0, a_lo = a_lo0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
          const int buf = tile_iter & 1;
          kwait(tempty_bar(buf), ((tile_iter >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem0 = tmem_base + buf * ACC_STRIDE;
          for (int kt = 0; kt < KT; ++kt) {
            const uint32_t d_tmem = d_tmem0 + ((DUAL && kt >= KT1) ? BN : 0);
            kwait(full_bar(stage), phase);
            if (consumer_fence) fence_proxy_async();
            tc_fence_after();
            const uint32_t b_lo = a_lo + (S::A_STAGE >> 4);
            const uint32_t first = (kt != 0 && !(DUAL && kt == KT1)) ? 1u : 0u;
            umma_i8(d_tmem, desc_hi | a_lo, desc_hi | b_lo, idesc, first);
            umma_i8(d_tmem, desc_hi | (a_lo + 2), desc_hi | (b_lo + 2), idesc, 1u);
            umma_commit(empty_bar(stage));
            if (kt == KT - 1) umma_commit(tfull_bar(buf));
            ++stage;
            a_lo += (S::STAGE >> 4);
            if (stage == STAGES) { stage = 0; phase ^= 1; a_lo = a_lo0; }
          }
        }
      }
    }
  } else {
    // =============================================================================== epilogue (EW warps)
    const int ew = warp - EPI_WARP0;             // 0..EW-1
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int half = ew >> 2;                    // column block of CW columns (2 halves, or 4 quarters when EW = 16)
    // slices: RES22 -> [0],[1] residual prefetch ring, [2] output; RES4x -> [0] residual in / output in place; RAW -> [0] output
    uint8_t* slice0 = smem + S::SLICES_OFF + ew * S::SLICE * S::SLICE_BUFS;
    uint8_t* yslice = slice0 + (S::SLICE_BUFS - 1) * S::SLICE;
    uint8_t* lowslice = smem + S::LOW_OFF + ew * S::LOW_SLICE;
    uint8_t* myy = yslice + lane * PITCH;
    uint8_t* mylow = lowslice + lane * S::LOW_PITCH;
    const int low_bits = IS_RES ? p.low_bits : (EPI == TC_EPI_REQ ? p.out_bits : 0);
    constexpr double kMagic = 6755399441055744.0, kOffS = 4503601774854144.0, kOffU = 4503599627370496.0;
    const double low_M = dyadic_to_double(p.low_m, p.low_e);
    const double res_M = dyadic_to_double(p.res_m, p.res_e);
    // unsigned operands (uint16 residual, post-ReLU sum): u * M + magic == fma(2^52 + u, M, magic - 2^52 * M) with a single
    // rounding of the same real number, as long as the folded constant is exact (e <= 51, checked below): saves one DADD each
    const double low_C = kMagic - kOffU * low_M;
    const double res_C = kMagic - kOffU * res_M;
    // post-ReLU values requantised with a ratio >= 0 are >= 0: with clamp [<= 0, 127] the clamp is the s8 saturation of cvt.pack
    const bool sat8 = IS_RES && p.sat_pack != 0 && p.low_bits == 8 && p.low_hi == 127 && p.low_lo <= 0;
    const int q_lo = (EPI == TC_EPI_REQ) ? (p.relu ? max(p.lo, 0) : p.lo) : p.low_lo;
    const int q_hi = (EPI == TC_EPI_REQ) ? p.hi : p.low_hi;
    constexpr int RES_CPR = CW * RES_ES / 16;    // 16-byte chunks per residual row
    int ymax = 0, ovf = 0;
    int bad = (IS_RES && !p.relu) ? 1 : 0;
    int cur_n0 = -1;
    uint32_t tile_iter = 0;
    auto ratio_ok = [](uint32_t m, int e) { return m == 0u || e >= (WIDE ? 11 : 31); };
    if (IS_RES && p.res_kind == 0) bad |= !ratio_ok(p.res_m, p.res_e) | (p.res_m != 0u && p.res_e > 51);
    if (low_bits && IS_RES) bad |= !dyadic_is_fast(p.low_m, p.low_e) | (p.low_m != 0u && p.low_e > 51);

    constexpr int RB = CW * 2;                  // RES22 tile row bytes (uint16)
    auto prefetch_residual_tma = [&](int tile, uint32_t k) {      // k-th tile of this CTA -> buffer k & 1
      if (lane == 0) {
        const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;   // m fastest: weights / constants change rarely
        mbar_arrive_expect_tx(res_bar(ew, k & 1), 32 * RB);
        tma_load_2d(smem_u32(slice0 + (k & 1) * S::SLICE), &maps.res, (n0 + half * CW) * 2, m0 + quarter * 32, res_bar(ew, k & 1));
      }
    };
    auto prefetch_residual = [&](int tile, uint8_t* dst) {
      if constexpr (IS_RES && !DUAL) {
        const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;   // m fastest: weights / constants change rarely
        const uint8_t* gres = reinterpret_cast<const uint8_t*>(p.res) + ((size_t)(m0 + quarter * 32) * p.Cout + n0 + half * CW) * RES_ES;
        const int rows_ok = p.M - (m0 + quarter * 32);
#pragma unroll
        for (int i = 0; i < RES_CPR; ++i) {       // 32 rows x RES_CPR chunks, a warp instruction covers whole rows
          const int id = lane + i * 32;
          const int rr = id / RES_CPR, j = id % RES_CPR;
          const bool v = rr < rows_ok;
          cp_async_16(smem_u32(dst + rr * PITCH + j * 16), v ? gres + (size_t)rr * p.Cout * RES_ES + j * 16 : gres, v ? 16 : 0);
        }
      }
      cp_async_commit();
    };
    // two values -> two uint16 halves with unsigned saturation (lo in the low half)
    auto pack2_sat_u16 = [](int lo, int hi) {
      uint32_t out;
      asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(out) : "r"(hi), "r"(lo));
      return out;
    };
    // pack 4 values into 4 bytes with signed-byte saturation (I2IP): byte i = sat_s8(v_i)
    auto pack4_sat_s8 = [](int a, int b, int c, int d) {
      uint32_t hi, out;
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(b), "r"(a), "r"(hi));
      return out;
    };
    // pack 4 clamped values into 4 bytes
    auto pack4 = [](int a, int b, int c, int d) { return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410); };

    if constexpr (EPI == TC_EPI_RES22) {
      if ((int)blockIdx.x < num_tiles) prefetch_residual_tma(blockIdx.x, 0);
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
      const int m0 = (tile % m_tiles) * BM, n0 = (tile / m_tiles) * BN;   // m fastest: weights / constants change rarely
      const int buf = tile_iter & 1;
      const int c0 = n0 + half * CW;             // first global channel of this warp
      uint8_t* rslice = slice0 + (EPI == TC_EPI_RES22 ? (tile_iter & 1) * S::SLICE : 0);

      // per-channel constants of this tile's channel block (shared by the 8 epilogue warps)
      if (n0 != cur_n0) {
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32));     // everyone finished reading the previous block
        for (int i = tid - EPI_WARP0 * 32; i < BN; i += EW * 32) {
          const hawq_chan ch = p.chan[n0 + i];
          sCst[i] = make_double2(kOffS - (double)ch.bias, dyadic_to_double(ch.m, ch.e));
          bad |= !ratio_ok(ch.m, ch.e) | (ch.bias >= (1 << 29)) | (ch.bias <= -(1 << 29));
          if constexpr (DUAL) {
            const hawq_chan rc = p.chan2[n0 + i];
            sCst2[i] = make_double2(kOffS - (double)rc.bias, dyadic_to_double(rc.m, rc.e));
            bad |= !ratio_ok(rc.m, rc.e) | (rc.bias >= (1 << 29)) | (rc.bias <= -(1 << 29));
          } else if constexpr (EPI >= TC_EPI_RES44) {
            if (p.res_kind == 1) {
              const hawq_chan rc = p.res_chan[n0 + i];
              sM1[i] = dyadic_to_double(rc.m, rc.e);
              bad |= !ratio_ok(rc.m, rc.e);
            } else {
              sM1[i] = res_M;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32));
        cur_n0 = n0;
      }

      if constexpr (EPI == TC_EPI_RES22) {         // this tile's residual was prefetched; start the next one
        const int nxt = tile + gridDim.x;
        if (nxt < num_tiles) prefetch_residual_tma(nxt, tile_iter + 1);
      } else if constexpr (IS_RES && !DUAL) {        // in-place variant (copy-out of the previous tile is synchronous)
        prefetch_residual(tile, rslice);
      }

      kwait(tfull_bar(buf), (tile_iter >> 1) & 1);
      tc_fence_after();
      if constexpr (EPI == TC_EPI_RES22) {
        kwait(res_bar(ew, tile_iter & 1), (tile_iter >> 1) & 1);     // residual tile landed (TMA)
        if (lane == 0) bulk_wait_read_all();                            // previous tile's TMA stores have read y / low tiles
      } else if constexpr (DUAL) {
        if (lane == 0) bulk_wait_read_all();                            // previous tile's TMA stores have read y / low tiles
      } else if constexpr (IS_RES) {
        cp_async_wait<0>();
      }
      __syncwarp();
      const uint8_t* myres = rslice + lane * PITCH;

#pragma unroll
      for (int cb = 0; cb < CW; cb += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * ACC_STRIDE + half * CW + cb, acc);
        tmem_ld_wait();
        const double2* cst = sCst + half * CW + cb;
        if constexpr (EPI == TC_EPI_REQ) {
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            int q[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const double2 cm = cst[j + k];
              const double d = __hiloint2double(0x43300000, acc[j + k] ^ 0x80000000) - cm.x;
              q[k] = clampi(__double2loint(__fma_rn(d, cm.y, kMagic)), q_lo, q_hi);
            }
            if (low_bits == 8) {
              *reinterpret_cast<uint4*>(mylow + cb + j) =
                  make_uint4(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7]), pack4(q[8], q[9], q[10], q[11]), pack4(q[12], q[13], q[14], q[15]));
            } else {
              *reinterpret_cast<uint2*>(mylow + ((cb + j) >> 1)) =
                  make_uint2(pack_nibbles8(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7])),
                             pack_nibbles8(pack4(q[8], q[9], q[10], q[11]), pack4(q[12], q[13], q[14], q[15])));
            }
          }
        } else if constexpr (EPI == TC_EPI_RAW) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            int4 o;   // bias = kOffS - Cb (exact); (acc + bias) + magic is exact, the low word is the int32 sum
            o.x = __double2loint((__hiloint2double(0x43300000, acc[j + 0] ^ 0x80000000) - cst[j + 0].x) + kMagic);
            o.y = __double2loint((__hiloint2double(0x43300000, acc[j + 1] ^ 0x80000000) - cst[j + 1].x) + kMagic);
            o.z = __double2loint((__hiloint2double(0x43300000, acc[j + 2] ^ 0x80000000) - cst[j + 2].x) + kMagic);
            o.w = __double2loint((__hiloint2double(0x43300000, acc[j + 3] ^ 0x80000000) - cst[j + 3].x) + kMagic);
            *reinterpret_cast<int4*>(myy + (cb + j) * 4) = o;
          }
        } else {   // RESIDUAL (relu = 1): groups of 8 channels (one 16-byte vector of uint16 residuals / two of int32)
          const double* m1 = sM1 + half * CW + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            uint32_t lw[4];
            uint32_t acc2[16];               // dual mode: identity-conv accumulator columns cb + j .. + 15
            if constexpr (DUAL) {
              tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * ACC_STRIDE + BN + half * CW + cb + j, acc2);
              tmem_ld_wait();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int jj = j + h * 8;
              int r[8];
              if constexpr (DUAL) {
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = (int)acc2[h * 8 + k];
              } else if constexpr (RES_ES == 2) {
                const uint4 pr = *reinterpret_cast<const uint4*>(rslice + tma_tile_off(RB, lane, (cb + jj) >> 3));
                r[0] = pr.x & 0xFFFF; r[1] = pr.x >> 16; r[2] = pr.y & 0xFFFF; r[3] = pr.y >> 16;
                r[4] = pr.z & 0xFFFF; r[5] = pr.z >> 16; r[6] = pr.w & 0xFFFF; r[7] = pr.w >> 16;
              } else {
                const int4 pa = *reinterpret_cast<const int4*>(myres + (cb + jj) * 4);
                const int4 pb = *reinterpret_cast<const int4*>(myres + (cb + jj) * 4 + 16);
                r[0] = pa.x; r[1] = pa.y; r[2] = pa.z; r[3] = pa.w; r[4] = pb.x; r[5] = pb.y; r[6] = pb.z; r[7] = pb.w;
              }
              int y[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const double2 cm = cst[jj + k];
                const double d = __hiloint2double(0x43300000, acc[jj + k] ^ 0x80000000) - cm.x;
                const double qv = __fma_rn(d, cm.y, kMagic);
                const int v = __double2loint(qv);
                double dr, mr;
                if constexpr (DUAL) {          // identity accumulator + its bias, per-channel ratio
                  const double2 c2 = sCst2[half * CW + cb + jj + k];
                  dr = __hiloint2double(0x43300000, r[k] ^ 0x80000000) - c2.x;
                  mr = c2.y;
                } else {
                  dr = (RES_ES == 2) ? __hiloint2double(0x43300000, r[k]) : (__hiloint2double(0x43300000, r[k] ^ 0x80000000) - kOffS);
                  mr = (EPI == TC_EPI_RES22) ? res_M : m1[jj + k];
                }
                const double qr = __fma_rn(dr, mr, (!DUAL && RES_ES == 2) ? res_C : kMagic);
                const int vr = __double2loint(qr);
                const int sum = v + vr;
                if constexpr (WIDE) {
                  ovf |= (__double2hiint(qv) + (int)((uint32_t)v >> 31)) ^ 0x43380000;
                  ovf |= (__double2hiint(qr) + (int)((uint32_t)vr >> 31)) ^ 0x43380000;
                  ovf |= ((v ^ sum) & (vr ^ sum)) >> 31;                   // the sum itself wrapped
                }
                y[k] = max(sum, 0);
              }
              if (low_bits) {
                int q[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)                                   // y >= 0: unsigned conversion, no sign fix-up
                  q[k] = __double2loint(__fma_rn(__hiloint2double(0x43300000, y[k]), low_M, low_C));
                if (sat8) {
                  lw[h * 2 + 0] = pack4_sat_s8(q[0], q[1], q[2], q[3]);
                  lw[h * 2 + 1] = pack4_sat_s8(q[4], q[5], q[6], q[7]);
                } else {
#pragma unroll
                  for (int k = 0; k < 8; ++k) q[k] = clampi(q[k], q_lo, q_hi);
                  if (low_bits == 8) {
                    lw[h * 2 + 0] = pack4(q[0], q[1], q[2], q[3]);
                    lw[h * 2 + 1] = pack4(q[4], q[5], q[6], q[7]);
                  } else {
                    lw[h] = pack_nibbles8(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7]));
                  }
                }
              }
              if constexpr (Y_ES == 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) ymax = max(ymax, y[k]);
                uint4 o;
                if (p.sat_pack) {      // one saturating pack per pair (clamps to [0, 65535])
                  o.x = pack2_sat_u16(y[0], y[1]); o.y = pack2_sat_u16(y[2], y[3]);
                  o.z = pack2_sat_u16(y[4], y[5]); o.w = pack2_sat_u16(y[6], y[7]);
                } else {
                  o.x = __byte_perm(min(y[0], 65535), min(y[1], 65535), 0x5410);
                  o.y = __byte_perm(min(y[2], 65535), min(y[3], 65535), 0x5410);
                  o.z = __byte_perm(min(y[4], 65535), min(y[5], 65535), 0x5410);
                  o.w = __byte_perm(min(y[6], 65535), min(y[7], 65535), 0x5410);
                }
                if constexpr (S::TMA_IO) *reinterpret_cast<uint4*>(yslice + tma_tile_off(RB, lane, (cb + jj) >> 3)) = o;
                else *reinterpret_cast<uint4*>(myy + (cb + jj) * 2) = o;
              } else {
                *reinterpret_cast<int4*>(myy + (cb + jj) * 4) = make_int4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<int4*>(myy + (cb + jj) * 4 + 16) = make_int4(y[4], y[5], y[6], y[7]);
              }
            }
            if constexpr (S::TMA_IO) {   // dense swizzled low tile (TMA box): 8-bit rows of CW bytes, 4-bit rows of CW/2 bytes
              if (low_bits == 8) *reinterpret_cast<uint4*>(lowslice + tma_tile_off(CW, lane, (cb + j) >> 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              else if (low_bits == 4) *reinterpret_cast<uint2*>(lowslice + tma_tile_off(CW / 2, lane, (cb + j) >> 5) + (((cb + j) >> 1) & 8)) = make_uint2(lw[0], lw[1]);
            } else {
              if (low_bits == 8) *reinterpret_cast<uint4*>(mylow + cb + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              else if (low_bits == 4) *reinterpret_cast<uint2*>(mylow + ((cb + j) >> 1)) = make_uint2(lw[0], lw[1]);
            }
          }
        }
      }
      // accumulator buffer fully read: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));

      const int rows_ok = p.M - (m0 + quarter * 32);
      if constexpr (S::TMA_IO) {
        // TMA tile stores: every lane publishes its writes to the async proxy, one lane issues two box stores (rows past M
        // are clipped by the tensor map); they drain while the next tile is computed
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&maps.y, c0 * 2, m0 + quarter * 32, smem_u32(yslice));
          if (low_bits == 8) tma_store_2d(&maps.low, c0, m0 + quarter * 32, smem_u32(lowslice));
          else if (low_bits == 4) tma_store_2d(&maps.low, c0 >> 1, m0 + quarter * 32, smem_u32(lowslice));
          bulk_commit();
        }
      } else {
      // coalesced copy-out of the staged outputs: a warp instruction writes whole rows (4 rows x 128 B for uint16 tiles);
      // all shared-memory reads are issued before the first global store
      if constexpr (Y_ES != 0) {
        constexpr int CPR = CW * Y_ES / 16;
        uint8_t* gy = reinterpret_cast<uint8_t*>(p.out) + ((size_t)(m0 + quarter * 32) * p.Cout + c0) * Y_ES;
        int4 v[CPR];
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int id = lane + i * 32;
          v[i] = *reinterpret_cast<const int4*>(yslice + (id / CPR) * PITCH + (id % CPR) * 16);
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int id = lane + i * 32;
          if (id / CPR < rows_ok) *reinterpret_cast<int4*>(gy + (size_t)(id / CPR) * p.Cout * Y_ES + (id % CPR) * 16) = v[i];
        }
      }
      if (low_bits == 8) {
        constexpr int CPR = CW / 16;
        uint8_t* gl = reinterpret_cast<uint8_t*>(IS_RES ? p.out_low : p.out) + (size_t)(m0 + quarter * 32) * p.Cout + c0;
        int4 v[CPR];
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int id = lane + i * 32;
          v[i] = *reinterpret_cast<const int4*>(lowslice + (id / CPR) * S::LOW_PITCH + (id % CPR) * 16);
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int id = lane + i * 32;
          if (id / CPR < rows_ok) *reinterpret_cast<int4*>(gl + (size_t)(id / CPR) * p.Cout + (id % CPR) * 16) = v[i];
        }
      } else if (low_bits == 4) {   // rows of CW / 2 packed bytes
        constexpr int CPR = CW / 32;
        uint8_t* gl = reinterpret_cast<uint8_t*>(IS_RES ? p.out_low : p.out) + (((size_t)(m0 + quarter * 32) * p.Cout + c0) >> 1);
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
          const int id = lane + i * 32;
          const int rr = id / CPR, j = id % CPR;
          if (rr < rows_ok)
            *reinterpret_cast<int4*>(gl + (((size_t)rr * p.Cout) >> 1) + j * 16) = *reinterpret_cast<const int4*>(lowslice + rr * S::LOW_PITCH + j * 16);
        }
      }
      __syncwarp();   // staging slices are rewritten by the next tile
      }
    }
    if constexpr (S::TMA_IO) {
      if (lane == 0) bulk_wait_all();
    }
    if constexpr (IS_RES) {
      cp_async_wait<0>();
      if (Y_ES == 2 && ymax > 65535) atomicOr(p.status, HAWQ_FLAG_RESIDUAL_OVERFLOW);
    }
    if (bad) atomicOr(p.status, HAWQ_FLAG_BAD_RATIO);
    if (ovf) atomicOr(p.status, HAWQ_FLAG_REQUANT_OVERFLOW);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// One-time re-tiling of OHWI int8 weights for the tcgen05 kernel: block (n_tile, k_tile) = BN rows x 64 bytes, stored
// contiguously with the SWIZZLE_64B pattern already applied, so a k-tile of B is one linear bulk copy.
__global__ void __launch_bounds__(256) retile_weights_kernel(const int8_t* __restrict__ w, int Cout, int K, int BN, int8_t* __restrict__ out) {
  const int kt_total = K / 64;
  const long long chunks = (long long)Cout * K / 16;
  for (long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x; id < chunks; id += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(id & 3);
    const long long rk = id >> 2;                 // (row, k-tile)
    const int kt = (int)(rk % kt_total);
    const int row = (int)(rk / kt_total);
    const int nt = row / BN, r = row % BN;
    const int4 v = *reinterpret_cast<const int4*>(w + (size_t)row * K + kt * 64 + c * 16);
    *reinterpret_cast<int4*>(out + ((size_t)nt * kt_total + kt) * BN * 64 + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
}

}  // namespace hawq

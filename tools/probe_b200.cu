// Hardware probes that decide the round-2 convolution design (results: profiles/r02/probe_b200.txt).
//   A. tcgen05.mma reading its A operand IN PLACE from a larger swizzled K-major matrix at an arbitrary ROW offset
//      (start address = base + off * row_bytes, not atom aligned), with and without the descriptor's base-offset field.
//   B. chip-wide L2 -> shared-memory fill rate with bulk copies (distinct vs shared source, cluster multicast).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/probe_b200 tools/probe_b200.cu && ./gpurun_out/probe_b200
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < 20000000u; ++i) if (try_wait(bar, parity)) return;
  __trap();
}

// ------------------------------------------------------------------------------------------------ A
// RB = row bytes = swizzle span (32 / 64 / 128).  A: 512 rows x RB in shared memory, address-swizzled (16-byte chunk index XOR
// address bits [7, 7 + log2(RB / 16))).  B: 64 rows x RB.  D[i][n] = sum_k A[off + i][k] * B[n][k], i < 128.
template <int RB>
__global__ void __launch_bounds__(128, 1) probe_desc(const int8_t* __restrict__ Ag, const int8_t* __restrict__ Bg, int off, int use_bo, int32_t* __restrict__ D) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  constexpr int CH = RB / 16;             // chunks per row
  uint8_t* sA = smem;                      // 512 rows
  uint8_t* sB = smem + 512 * RB;           // 64 rows (1024-aligned: 512 * RB is a multiple of 1024)
  for (int id = threadIdx.x; id < 512 * CH; id += 128) {
    const int r = id / CH, j = id % CH;
    const uint32_t lin = (uint32_t)(r * RB);
    const uint32_t x = (lin >> 7) & (CH - 1);
    *reinterpret_cast<int4*>(sA + lin + ((j ^ x) << 4)) = *reinterpret_cast<const int4*>(Ag + (size_t)r * RB + j * 16);
  }
  for (int id = threadIdx.x; id < 64 * CH; id += 128) {
    const int r = id / CH, j = id % CH;
    const uint32_t lin = (uint32_t)(r * RB);
    const uint32_t x = (lin >> 7) & (CH - 1);
    *reinterpret_cast<int4*>(sB + lin + ((j ^ x) << 4)) = *reinterpret_cast<const int4*>(Bg + (size_t)r * RB + j * 16);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t layout = RB == 128 ? 2 : RB == 64 ? 4 : 6;
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)off * RB, b_addr = smem_u32(sB);
    auto desc = [&](uint32_t addr, uint32_t bo) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((8 * RB) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(bo & 7) << 49) | (layout << 61);
    };
    const uint32_t bo = use_bo ? ((a_addr >> 7) & 7) : 0;
    for (int k = 0; k < RB / 32; ++k) {
      const uint64_t ad = desc(a_addr + k * 32, bo), bd = desc(b_addr + k * 32, 0);
      const uint32_t acc = k != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  wait_bar(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 16; ++k) D[(warp * 32 + lane) * 64 + c + k] = (int32_t)v[k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

template <int RB>
void run_desc() {
  std::vector<int8_t> A(512 * RB), B(64 * RB);
  srand(1);
  for (auto& v : A) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : B) v = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB; int32_t* dD;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  const int smem = (512 + 64) * RB + 1024;
  CK(cudaFuncSetAttribute(probe_desc<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int offs[] = {0, 8, 1, 2, 3, 4, 5, 7, 13, 30, 58, 59, 117, 118, 200};
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int off : offs) {
      CK(cudaMemset(dD, 0xFF, 128 * 64 * 4));
      probe_desc<RB><<<1, 128, smem>>>(dA, dB, off, use_bo, dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("A: RB=%d off=%d bo=%d: %s\n", RB, off, use_bo, cudaGetErrorString(e)); exit(1); }
      std::vector<int32_t> D(128 * 64);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0, bad_rows = 0;
      for (int i = 0; i < 128; ++i) {
        int rb = 0;
        for (int n = 0; n < 64; ++n) {
          int ref = 0;
          for (int k = 0; k < RB; ++k) ref += (int)A[(off + i) * RB + k] * (int)B[n * RB + k];
          if (ref != D[i * 64 + n]) { ++bad; rb = 1; }
        }
        bad_rows += rb;
      }
      printf("A: swizzle %3dB row-offset %3d base_offset_field=%s : %s (%d wrong values in %d rows)\n", RB, off, use_bo ? "auto" : "0   ", bad ? "MISMATCH" : "exact", bad, bad_rows);
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// ------------------------------------------------------------------------------------------------ B
// Every CTA streams `iters` chunks of CHUNK bytes from global memory (L2 resident after warm-up) into a STAGES-deep ring with
// cp.async.bulk; a consumer thread only waits and releases.  mode 0: CTA-distinct chunks walking a `span`-byte region;
// mode 1: all CTAs read the same sequence of chunks (weights-like);  CL > 1: mode 1 with cluster multicast (one CTA issues 1/CL of the bytes
// to all CTAs of the cluster).
template <int CL, int STAGES = 8>
__global__ void __launch_bounds__(64, 1) probe_l2(const uint8_t* __restrict__ src, size_t span, int chunk, int iters, int mode, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  uint32_t rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(CL));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); }
  const size_t nchunks = span / chunk;
  if (threadIdx.x == 0) {            // producer
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      wait_bar(smem_u32(&empty[s]), ((it / STAGES) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(chunk) : "memory");
      const size_t c = mode == 0 ? ((size_t)blockIdx.x + (size_t)it * gridDim.x) % nchunks : (size_t)it % nchunks;
      if (CL == 1) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + s * chunk)), "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(&full[s])) : "memory");
      } else {
        const int part = chunk / CL;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(smem + s * chunk + rank * part)), "l"(src + c * chunk + rank * part), "r"(part), "r"(smem_u32(&full[s])), "h"((uint16_t)((1 << CL) - 1)) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {    // consumer
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      wait_bar(smem_u32(&full[s]), (it / STAGES) & 1);
      if (CL == 1) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      } else {
        for (uint32_t r = 0; r < CL; ++r) {     // release the stage in every CTA of the cluster (each one multicasts into all)
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[s])), "r"(r));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      }
    }
  }
  __syncthreads();
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;"); }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1;
}

// latency of ONE bulk copy (issue -> mbarrier completes), idle chip (grid 1) or every SM doing the same (grid 148)
__global__ void __launch_bounds__(32, 1) probe_lat(const uint8_t* __restrict__ src, int chunk, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    long long tot = 0;
    for (int i = 0; i < reps; ++i) {
      const long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem)), "l"(src + ((size_t)blockIdx.x * 64 + i) * chunk), "r"(chunk), "r"(smem_u32(&bar)) : "memory");
      while (!try_wait(smem_u32(&bar), i & 1)) {}
      tot += clock64() - t0;
    }
    if (blockIdx.x == 0) out[0] = tot / reps;
  }
}
void run_lat(const uint8_t* src, int chunk, int grid) {
  long long* d; CK(cudaMalloc(&d, 8));
  CK(cudaFuncSetAttribute(probe_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk + 1024));
  probe_lat<<<grid, 32, chunk + 1024>>>(src, chunk, 64, d);
  CK(cudaDeviceSynchronize());
  probe_lat<<<grid, 32, chunk + 1024>>>(src, chunk, 64, d);
  CK(cudaDeviceSynchronize());
  long long cyc; CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
  printf("C: one bulk copy of %5d B at a time, %3d CTAs: %lld cycles issue -> complete (L2-resident source)\n", chunk, grid, cyc);
  cudaFree(d);
}

template <int CL, int STAGES = 8>
void run_l2(const uint8_t* src, size_t span, int chunk, int mode, const char* what, int grid) {
  unsigned long long* d; CK(cudaMalloc(&d, 8));
  const int smem = STAGES * chunk + 1024;
  CK(cudaFuncSetAttribute((probe_l2<CL, STAGES>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int iters = 4000;
  CK(cudaLaunchKernelEx(&cfg, (probe_l2<CL, STAGES>), src, span, chunk, 400, mode, d));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  CK(cudaLaunchKernelEx(&cfg, (probe_l2<CL, STAGES>), src, span, chunk, iters, mode, d));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)grid * iters * chunk;
  printf("B: %-44s grid %3d cluster %d stages %2d chunk %5d B span %4zu MB: %.2f TB/s into shared memory (%.1f GB/s per SM)\n", what, grid, CL, STAGES, chunk, span >> 20, bytes / (ms * 1e-3) / 1e12,
         bytes / (ms * 1e-3) / 1e9 / grid);
  cudaFree(d);
}


// ------------------------------------------------------------------------------------------------ D
// 2-D tensor TMA: boxes of `rows` x INNER bytes (SWIZZLE matching INNER) streamed into an 8-deep ring by NPROD producer threads
// (one per warp, round robin over the stages).  Answers: what does a 64-byte inner row cost against a 128-byte one, and does a
// second issuing warp raise the per-SM copy rate?
#include <cuda.h>
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int NPROD>
__global__ void __launch_bounds__(32 * (NPROD + 1), 1) probe_tma(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int total_rows, int iters, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  constexpr int STAGES = 8;
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nboxes = total_rows / box_rows;
  const int stage_bytes = (box_bytes + 1023) / 1024 * 1024;
  if (warp < NPROD && lane == 0) {
    for (int it = warp; it < iters; it += NPROD) {
      const int s = it % STAGES;
      wait_bar(smem_u32(&empty[s]), ((it / STAGES) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(box_bytes) : "memory");
      const int b = (blockIdx.x + it * gridDim.x) % nboxes;
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(smem + s * stage_bytes)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(0), "r"(b * box_rows), "r"(smem_u32(&full[s])) : "memory");
    }
  } else if (warp == NPROD && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      wait_bar(smem_u32(&full[s]), (it / STAGES) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1;
}
template <int NPROD>
void run_tma(uint8_t* src, int inner, int rows) {
  static encode_tiled_fn enc = [] {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    return reinterpret_cast<encode_tiled_fn>(ptr);
  }();
  const int total_rows = (32 << 20) / inner;
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)total_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)inner};
  const cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (!enc || enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("D: encode failed\n"); return; }
  unsigned long long* d; CK(cudaMalloc(&d, 8));
  const int box_bytes = inner * rows;
  const int smem = 8 * ((box_bytes + 1023) / 1024 * 1024) + 1024;
  CK(cudaFuncSetAttribute(probe_tma<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4000;
  probe_tma<NPROD><<<148, 32 * (NPROD + 1), smem>>>(map, box_bytes, rows, total_rows, 400, d);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_tma<NPROD><<<148, 32 * (NPROD + 1), smem>>>(map, box_bytes, rows, total_rows, iters, d);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 148.0 * iters * box_bytes;
  printf("D: tensor TMA box %3d B x %3d rows (%5d B), %d issuing warp(s): %.2f TB/s, %.0f ns per box per SM, %.1f ns per box row\n", inner, rows, box_bytes, NPROD,
         bytes / (ms * 1e-3) / 1e12, ms * 1e6 / iters, ms * 1e6 / iters / rows);
  cudaFree(d);
}

// ------------------------------------------------------------------------------------------------ E
// 4-D boxes with halo (what conv_halo loads): tensor {C bytes, W, H, N}, box {64, W + 2, R + 2, 1} at (0, -1, y0 - 1, n).
__global__ void __launch_bounds__(64, 1) probe_tma4(const __grid_constant__ CUtensorMap map, int box_bytes, int H, int R, int N, int iters, int nbuf, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int stage_bytes = (box_bytes + 1023) / 1024 * 1024;
  const int tiles_per_img = (H + R - 1) / R;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % nbuf;
      wait_bar(smem_u32(&empty[s]), ((it / nbuf) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(box_bytes) : "memory");
      const int mt = (blockIdx.x + it * gridDim.x) % (tiles_per_img * N);
      const int n = mt / tiles_per_img, y0 = (mt % tiles_per_img) * R;
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(smem_u32(smem + s * stage_bytes)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(0), "r"(-1), "r"(y0 - 1), "r"(n), "r"(smem_u32(&full[s])) : "memory");
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % nbuf;
      wait_bar(smem_u32(&full[s]), (it / nbuf) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
}
void run_tma4(uint8_t* src, int C, int W, int H, int N, int R, int nbuf, int inner = 64) {
  static encode_tiled_fn enc = [] {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    return reinterpret_cast<encode_tiled_fn>(ptr);
  }();
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C, (cuuint64_t)C * W, (cuuint64_t)C * W * H};
  const cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)(W + 2), (cuuint32_t)(R + 2), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!enc || enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("E: encode failed\n"); return; }
  long long* d; CK(cudaMalloc(&d, 8));
  const int box_bytes = inner * (W + 2) * (R + 2);
  const int smem = nbuf * ((box_bytes + 1023) / 1024 * 1024) + 1024;
  CK(cudaFuncSetAttribute(probe_tma4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  probe_tma4<<<148, 64, smem>>>(map, box_bytes, H, R, N, 200, nbuf, d);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_tma4<<<148, 64, smem>>>(map, box_bytes, H, R, N, iters, nbuf, d);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("E: 4-D halo box {%d B, %d, %d, 1} = %5d B over {%d, %d, %d, %d}, %d buffers: %.0f ns per box per SM, %.2f TB/s\n", inner, W + 2, R + 2, box_bytes, C, W, H, N, nbuf,
         ms * 1e6 / iters, 148.0 * iters * box_bytes / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  if (getenv("PROBE_DESC")) { run_desc<64>(); run_desc<128>(); run_desc<32>(); }
  uint8_t* src; const size_t cap = 512ull << 20;
  CK(cudaMalloc(&src, cap)); CK(cudaMemset(src, 1, cap));
  if (!getenv("PROBE_E_ONLY")) {
  run_tma<1>(src, 64, 128); run_tma<1>(src, 64, 232); run_tma<1>(src, 128, 64); run_tma<1>(src, 128, 116);
  run_tma<1>(src, 32, 232); run_tma<1>(src, 64, 32); run_tma<1>(src, 128, 16);
  run_tma<2>(src, 64, 128); run_tma<2>(src, 64, 232); run_tma<4>(src, 64, 128); run_tma<2>(src, 128, 64);
  }
  run_tma4(src, 64, 56, 56, 128, 2, 4); run_tma4(src, 64, 56, 56, 128, 2, 8); run_tma4(src, 64, 56, 56, 128, 2, 1);
  run_tma4(src, 128, 28, 28, 128, 4, 4); run_tma4(src, 128, 28, 28, 128, 4, 4, 128); run_tma4(src, 256, 14, 14, 128, 8, 4); run_tma4(src, 256, 14, 14, 128, 8, 4, 128);
  run_tma4(src, 64, 56, 56, 128, 0, 4); run_tma4(src, 64, 56, 56, 128, 6, 2);
  if (getenv("PROBE_TMA_ONLY")) return 0;
  for (int chunk : {8192, 16384, 24576}) {
    run_l2<1>(src, 32ull << 20, chunk, 0, "distinct chunks, L2-resident region", 148);
    run_l2<1>(src, 512ull << 20, chunk, 0, "distinct chunks, HBM-sized region", 148);
    run_l2<1>(src, 2ull << 20, chunk, 1, "all CTAs read the same chunks (unicast)", 148);
  }
  run_l2<2>(src, 2ull << 20, 16384, 1, "same chunks, multicast pairs", 148);
  run_l2<4>(src, 2ull << 20, 16384, 1, "same chunks, multicast clusters of 4", 148);
  run_l2<1>(src, 32ull << 20, 16384, 0, "distinct chunks, L2-resident, half the SMs", 74);
  run_l2<1, 2>(src, 32ull << 20, 8192, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 4>(src, 32ull << 20, 8192, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 16>(src, 32ull << 20, 8192, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 24>(src, 32ull << 20, 8192, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 16>(src, 32ull << 20, 4096, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 4>(src, 32ull << 20, 32768, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 6>(src, 32ull << 20, 32768, 0, "distinct chunks, L2-resident", 148);
  run_l2<1, 16>(src, 512ull << 20, 8192, 0, "distinct chunks, HBM-sized", 148);
  for (int chunk : {1024, 8192, 16384, 32768}) { run_lat(src, chunk, 1); run_lat(src, chunk, 148); }
  return 0;
}

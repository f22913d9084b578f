// Issue-rate probe of tcgen05.mma kind::i8 (M = 128) with both operands in shared memory: cycles per instruction as a function of
// N, of the operand layout (64-byte / 128-byte swizzled K-major rows) and of the row offset of the A descriptor (the 3x3 kernel
// reads tap-shifted views of one patch, so its A descriptors start at arbitrary rows).  Operand contents are irrelevant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hawq_b200/csrc -I include -o /tmp/probe_mma tools/probe_mma.cu && /tmp/probe_mma
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

using namespace hawq;

__device__ __forceinline__ uint64_t desc_for(uint32_t addr, int sw_bytes) {
  const uint64_t sbo = (uint64_t)(8 * sw_bytes) >> 4;                  // 8 rows of sw_bytes
  const uint64_t mode = sw_bytes == 128 ? 2 : sw_bytes == 64 ? 4 : 6;  // SWIZZLE_128B / 64B / 32B
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | (sbo << 32) | ((uint64_t)1 << 46) | (mode << 61);
}

// mode: which A start addresses the instruction stream cycles through
//   0: one aligned tile, k steps of 32 bytes inside the row        1: rows shifted by `shift` rows per instruction (tap-shifted views)
__global__ void __launch_bounds__(64, 1) probe_mma(int n, int sw_bytes, int shift_rows, int iters, int epi_lds, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 96 * 1024;
  const uint32_t idesc = umma_idesc_i8(128, n, true);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int kstep = i % (sw_bytes / 32);
        const uint32_t a = a_base + (uint32_t)((i % 9) * shift_rows * sw_bytes) + kstep * 32;
        const uint32_t b = b_base + (uint32_t)((i % 4) * n * sw_bytes) + kstep * 32;
        umma_i8(tmem_base, desc_for(a, sw_bytes), desc_for(b, sw_bytes), idesc, i > 0);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      t1 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; }
    }
  } else if (epi_lds) {
    // a second warp hammering shared memory with broadcast LDS.128 while the MMAs run: how much does it slow them / itself
    const uint4* p = reinterpret_cast<const uint4*>(smem + 200 * 1024);
    uint32_t acc = 0;
    const long long s0 = clock64();
    for (int i = 0; i < epi_lds; ++i) {
      const uint4 v = p[(i & 63)];
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    const long long s1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[1] = s1 - s0; out[2] = acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem_base); }
}

// 16 extra warps run one micro-loop (registers only unless stated) while warp 0 streams MMAs: which resource do epilogue warps
// lose to the tensor pipe?  work: 0 nothing, 1 DFMA (8 independent chains), 2 broadcast LDS.128, 3 tcgen05.ld x16 + wait,
// 4 IMAD.WIDE (8 chains), 5 STS.128 (conflict-free), 6 FFMA (8 chains)
__global__ void __launch_bounds__(17 * 32, 1) probe_mix(int n, int mma_iters, int work, int work_iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 96 * 1024;
  const uint32_t idesc = umma_idesc_i8(128, n, true);
  if (warp == 0) {
    if (elect_one()) {
      const long long t0 = clock64();
      for (int i = 0; i < mma_iters; ++i) {
        const int kstep = i & 1;
        const uint32_t a = a_base + (uint32_t)((i % 9) * 3 * 64) + kstep * 32;
        const uint32_t b = b_base + (uint32_t)((i % 4) * n * 64) + kstep * 32;
        umma_i8(tmem_base, desc_for(a, 64), desc_for(b, 64), idesc, i > 0);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (work) {
    const int ew = warp - 1, quarter = warp & 3;
    const long long s0 = clock64();
    uint32_t sink = 0;
    if (work == 1) {
      double a[8];
      for (int j = 0; j < 8; ++j) a[j] = 1.0 + j + lane;
      const double m = 0.999999 + 1e-9 * lane, c = 1e-3;
      for (int i = 0; i < work_iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __fma_rn(a[j], m, c);
      for (int j = 0; j < 8; ++j) sink ^= (uint32_t)__double2loint(a[j]);
    } else if (work == 2) {
      const uint4* p = reinterpret_cast<const uint4*>(smem + 200 * 1024) + ew * 8;
      for (int i = 0; i < work_iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const uint4 v = p[(i + j) & 7]; sink ^= v.x + v.y + v.z + v.w; }
    } else if (work == 3) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256 + (ew >> 2) * 64;
      for (int i = 0; i < work_iters; ++i) {
        uint32_t v[16];
        tmem_ld16(taddr + (i & 3) * 16, v);
        tmem_ld_wait();
        sink ^= v[0] + v[5] + v[15];
      }
    } else if (work == 4) {
      long long a[8];
      for (int j = 0; j < 8; ++j) a[j] = j + lane;
      const int m = 0x6789ABCD + lane;
      for (int i = 0; i < work_iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = (long long)(int)(a[j] >> 7) * m + a[j];
      for (int j = 0; j < 8; ++j) sink ^= (uint32_t)a[j];
    } else if (work == 5) {
      uint4* p = reinterpret_cast<uint4*>(smem + 200 * 1024) + ew * 32 + lane;
      for (int i = 0; i < work_iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) p[0] = make_uint4(i, j, lane, sink);
    } else if (work == 6) {
      float a[8];
      for (int j = 0; j < 8; ++j) a[j] = 1.0f + j + lane;
      const float m = 0.99999f, c = 1e-3f;
      for (int i = 0; i < work_iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __fmaf_rn(a[j], m, c);
      for (int j = 0; j < 8; ++j) sink ^= __float_as_uint(a[j]);
    }
    const long long s1 = clock64();
    if (blockIdx.x == 0 && warp == 1 && lane == 0) { out[1] = s1 - s0; out[2] = sink; }
    if (sink == 0x12345678u) out[3] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

static void run_mix() {
  long long* out;
  cudaMalloc(&out, 64);
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(probe_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"none", "DFMA x8", "LDS.128 bcast x8", "LDTM x16+wait", "IMAD.WIDE x8", "STS.128 x8", "FFMA x8"};
  printf("\n16 warps of side work against a stream of M=128 MMAs (all 148 SMs); side work: iterations of 8 operations per thread\n");
  printf("%4s %-18s | %12s %14s | %14s %12s\n", "N", "side work", "clk/mma", "(alone)", "clk/iter", "(mma idle)");
  for (int n : {64, 128, 256}) {
    double alone_mma = 0;
    for (int work = 0; work <= 6; ++work) {
      const int witers = 2000, miters = 2048;
      long long h[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      cudaMemset(out, 0, 64);
      probe_mix<<<148, 17 * 32, smem>>>(n, miters, work, witers, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("mix N=%d work=%d failed: %s\n", n, work, cudaGetErrorString(cudaGetLastError())); return; }
      cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
      if (work) {
        cudaMemset(out, 0, 64);
        probe_mix<<<148, 17 * 32, smem>>>(n, 1, work, witers, out);       // side work with an idle tensor pipe
        cudaDeviceSynchronize();
        cudaMemcpy(hi, out, 24, cudaMemcpyDeviceToHost);
      } else {
        alone_mma = (double)h[0] / miters;
      }
      printf("%4d %-18s | %12.1f %14.1f | %14.1f %12.1f\n", n, names[work], (double)h[0] / miters, alone_mma, work ? (double)h[1] / witers : 0.0,
             work ? (double)hi[1] / witers : 0.0);
    }
  }
}

int main() {
  if (getenv("PROBE_MIX")) { run_mix(); return 0; }
  long long* out;
  cudaMalloc(&out, 64);
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 1024;
  printf("tcgen05.mma kind::i8 M=128, both operands in shared memory, %d back-to-back instructions on every SM (148 CTAs)\n", iters);
  printf("%4s %5s %6s %8s | %10s %10s %10s\n", "N", "swz", "shift", "lds", "clk/mma", "MAC/clk", "lds clk/ld");
  for (int sw : {64, 128})
    for (int n : {64, 128, 256})
      for (int shift : {0, 3})
        for (int lds : {0, 4096}) {
          if (sw == 128 && n == 256) continue;      // (the B tile would leave the 220 KB buffer)
          cudaMemset(out, 0, 64);
          probe_mma<<<148, 64, smem>>>(n, sw, shift, iters, lds, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("N=%d sw=%d shift=%d: %s\n", n, sw, shift, cudaGetErrorString(e)); return 1; }
          long long h[3];
          cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
          const double c = (double)h[0] / iters;
          printf("%4d %5d %6d %8d | %10.1f %10.0f %10.1f\n", n, sw, shift, lds, c, 128.0 * n * 32 / c, lds ? (double)h[1] / lds : 0.0);
        }
  return 0;
}

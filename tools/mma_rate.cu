// Microbenchmark: issue rate of tcgen05.mma kind::i8 vs kind::f8f6f4 (e4m3) with shared-memory operands, no loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && ./gpurun_out/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)32 << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
template <int KIND>  // 0 = i8, 1 = f8f6f4 (e4m3 x e4m3 -> f32)
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < (128 + N) * 64; i += 128) smem[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  // idesc: i8: c=S32(2), a,b signed(1); f8: c=F32(1), a,b E4M3(0)
  const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint64_t a = desc_sw64(smem_u32(smem)), b = desc_sw64(smem_u32(smem) + 128 * 64);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mma<KIND>(tm, a, b, idesc, i != 0);
      mma<KIND>(tm, a + 2, b + 2, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}
template <int KIND, int N>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 4096;
  auto fn = k<KIND, N>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + N) * 64 + 1024);
  fn<<<148, 128, (128 + N) * 64 + 1024>>>(16, d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  fn<<<148, 128, (128 + N) * 64 + 1024>>>(iters, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double macs = 2.0 * iters * 128.0 * N * 32.0;
  printf("%-22s N=%3d: %s  %.1f cycles per MMA (M128,K32), %.0f MAC/clk/SM, chip %.2f POPS (event %.3f ms)\n", name, N, cudaGetErrorString(err),
         (double)cyc / (2.0 * iters), macs / cyc, 148 * macs * 2 / (ms * 1e-3) / 1e15, ms);
}
int main() {
  run<0, 64>("kind::i8");  run<0, 128>("kind::i8");  run<0, 256>("kind::i8");
  run<1, 64>("kind::f8f6f4 e4m3");  run<1, 128>("kind::f8f6f4 e4m3");  run<1, 256>("kind::f8f6f4 e4m3");
  return 0;
}

// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16, SS operands, 128-byte swizzle) from one thread per CTA,
// no TMA involved (smem content is whatever is there).  Prints cycles per MMA for a few shapes / accumulator patterns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0; d |= (uint64_t)((addr & 0x3FFFF) >> 4); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
template <int N, int NACC, int KSTEP>
__global__ void __launch_bounds__(128) rate(long long* out, int nmma) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 16384 * 2;
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const int k = KSTEP ? (i & 3) : 0;       // advance inside a 64-wide K block like the real kernel, or re-read the same 16
      const int st = (i >> 2) % 3;             // 3 smem stages of 32 KB like the real ring
      uint64_t ad = make_desc(a + st * 49152 + k * 32, 16, 1024), bd = make_desc(b + st * 49152 + k * 32, 16, 1024);
      uint32_t d = tm + (uint32_t)((i / 4) % NACC) * N, acc = 1;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc(128, N)), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    long long t1 = clock64();
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int N, int NACC, int KSTEP> void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16); const int nmma = 2048;
  auto k = rate<N, NACC, KSTEP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<grid, 128, 200 * 1024>>>(d, nmma); cudaDeviceSynchronize();
  k<<<grid, 128, 200 * 1024>>>(d, nmma); cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s grid %3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %d)  %s\n", name, grid, (double)h[0] / nmma, (double)h[1] / nmma, 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<128, 1, 1>("M128 N128 1 accumulator, k-advance", 1);
  run<128, 1, 1>("M128 N128 1 accumulator, k-advance", 148);
  run<128, 2, 1>("M128 N128 2 accumulators", 148);
  run<128, 1, 0>("M128 N128 same 16-K slice", 148);
  run<256, 1, 1>("M128 N256 1 accumulator", 148);
  run<64, 1, 1>("M128 N64 1 accumulator", 148);
  return 0;
}

// Micro-benchmark: what limits a single CTA's tcgen05.mma stream once it is throttled by mbarrier hand-shakes like the real
// tap-GEMM pipeline?  One CTA per SM, no TMA (operands are whatever is in smem).  Variants:
//   0: 4 MMAs + commit per step, nobody waits                      (pure issue rate with commits)
//   1: MMA thread waits on the commit barrier of step i-S itself   (completion latency, S stages)
//   2: a second thread (other warp) relays empty[s] -> full[s], the MMA thread waits on full[s]   (the real structure)
//   3: like 2, plus 128 more threads polling an unrelated barrier   (epilogue warps waiting for the accumulator)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0; d |= (uint64_t)((addr & 0x3FFFF) >> 4); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void bar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }

__device__ __forceinline__ void bar_wait_test(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
// two issuer threads (different warps) alternate steps; a relay thread turns empty[s] into full[s]
template <int N, int S, int MPS, int TESTWAIT, int NISS>
__global__ void __launch_bounds__(256) pipe2(long long* out, int steps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], fin; __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { bar_init(&full[i], 1); bar_init(&empty[i], 1); } bar_init(&fin, NISS); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  constexpr int STAGE = (MPS / 4) * (16384 + N * 128);
  const int who = threadIdx.x == 0 ? 0 : (threadIdx.x == 64 ? 1 : -1);
  if (who >= 0 && who < NISS) {
    long long t0 = clock64();
    for (int i = who; i < steps; i += NISS) {
      const int s = i % S; const uint32_t ph = (uint32_t)(i / S) & 1u;
      if (TESTWAIT) bar_wait_test(&full[s], ph); else bar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a = smem_u32(smem) + s * STAGE, b = a + (MPS / 4) * 16384;
#pragma unroll
      for (int k = 0; k < MPS; ++k) {
        uint64_t ad = make_desc(a + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), bd = make_desc(b + (k >> 2) * N * 128 + (k & 3) * 32, 16, 1024);
        uint32_t acc = ((i - who) | k) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm + (uint32_t)who * (N < 256 ? N : 0)), "l"(ad), "l"(bd), "r"(idesc(128, N)), "r"(acc));
      }
      commit(&empty[s]);
    }
    commit(&fin);
    long long t1 = clock64();
    bar_wait(&fin, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && who == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < steps; ++i) {
      bar_wait(&empty[s], ph ^ 1u);
      bar_arrive(&full[s]);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}
template <int N, int S, int MPS, int TESTWAIT, int NISS> void run2(const char* name, int per_sm = 1) {
  long long* d; cudaMalloc(&d, 16); const int steps = 512;
  auto k = pipe2<N, S, MPS, TESTWAIT, NISS>;
  const int smem = per_sm == 1 ? 210 * 1024 : 100 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<148 * per_sm, 256, smem>>>(d, steps); cudaDeviceSynchronize();
  k<<<148 * per_sm, 256, smem>>>(d, steps); cudaError_t e = cudaDeviceSynchronize();
  if (per_sm > 1) printf("[%d CTAs/SM] ", per_sm);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-64s N=%3d S=%d MMA/step=%d: complete %7.1f cyc/step (MMA ideal %d)  %s\n", name, N, S, MPS, (double)h[1] / steps, MPS * (N < 128 ? 64 : N / 2), cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int S, int VAR>
__global__ void __launch_bounds__(256) pipe(long long* out, int steps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], fin, never; __shared__ uint32_t slot;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { bar_init(&full[i], 1); bar_init(&empty[i], 1); } bar_init(&fin, 1); bar_init(&never, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  constexpr int STAGE = 16384 + N * 128;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < steps; ++i) {
      if (VAR == 1 && i >= S) bar_wait(&empty[s], ph ^ 1u);        // own completion of step i-S (phase ph^1 completed)
      if (VAR >= 2) { bar_wait(&full[s], ph); asm volatile("tcgen05.fence::after_thread_sync;"); }
      const uint32_t a = smem_u32(smem) + s * STAGE, b = a + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = make_desc(a + k * 32, 16, 1024), bd = make_desc(b + k * 32, 16, 1024);
        uint32_t acc = (i | k) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc(128, N)), "r"(acc));
      }
      commit(&empty[s]);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    commit(&fin);
    long long t1 = clock64();
    bar_wait(&fin, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    bar_arrive(&never);
  } else if (threadIdx.x == 32 && VAR >= 2) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < steps; ++i) {
      bar_wait(&empty[s], ph ^ 1u);
      bar_arrive(&full[s]);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x >= 64 && threadIdx.x < 192 && VAR == 3) {
    bar_wait(&never, 0);
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}
template <int N, int S, int VAR> void run(const char* name) {
  long long* d; cudaMalloc(&d, 16); const int steps = 512;
  auto k = pipe<N, S, VAR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  k<<<148, 256, 210 * 1024>>>(d, steps); cudaDeviceSynchronize();
  k<<<148, 256, 210 * 1024>>>(d, steps); cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-64s N=%3d S=%d: issue %7.1f cyc/step, complete %7.1f cyc/step (MMA ideal %d)  %s\n", name, N, S, (double)h[0] / steps, (double)h[1] / steps, 4 * (N < 128 ? 64 : N / 2), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<128, 3, 0>("0: 4 MMA + commit per step, no waits");
  run<128, 3, 1>("1: self-wait on the commit of step i-S");
  run<128, 6, 1>("1: self-wait on the commit of step i-S");
  run<128, 3, 2>("2: relay thread empty->full, MMA thread waits full");
  run<128, 6, 2>("2: relay thread empty->full, MMA thread waits full");
  run<128, 3, 3>("3: relay + 128 threads polling another barrier");
  run<128, 6, 3>("3: relay + 128 threads polling another barrier");
  run<64, 4, 2>("2: relay, N=64");
  run<256, 4, 2>("2: relay, N=256");
  run2<128, 3, 4, 0, 1>("relay, 1 issuer, try_wait (= variant 2)");
  run2<128, 3, 4, 1, 1>("relay, 1 issuer, test_wait");
  run2<128, 3, 8, 0, 1>("relay, 1 issuer, 8 MMAs per step (BK=128)");
  run2<128, 4, 4, 0, 2>("relay, 2 alternating issuers");
  run2<128, 4, 8, 0, 2>("relay, 2 alternating issuers, 8 MMAs per step");
  run2<64, 4, 8, 0, 1>("relay, 1 issuer, N=64, 8 MMAs per step");
  run2<64, 4, 8, 0, 2>("relay, 2 issuers, N=64, 8 MMAs per step");
  run2<128, 3, 4, 0, 1>("relay, 1 issuer", 2);
  run2<128, 3, 4, 0, 2>("relay, 2 issuers", 2);
  run2<64, 4, 4, 0, 1>("relay, 1 issuer, N=64", 2);
  run2<64, 4, 4, 0, 2>("relay, 2 issuers, N=64", 2);
  return 0;
}

// Shared device/host helpers for the stcgan_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include "../../include/stcgan_b200.h"

namespace stcgan {

// ---------------------------------------------------------------------------------------------
// launch accounting + error plumbing
// ---------------------------------------------------------------------------------------------
// (a relaxed atomic: launches may be issued from several host threads -- one per device under nn.DataParallel-style callers;
// it is a statistics counter only, no kernel or host logic depends on its value)
extern std::atomic<int64_t> g_launches;

inline int finish_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// Programmatic dependent launch (PDL): every kernel of this library is launched with the programmatic-stream-serialization
// attribute, signals `launch_dependents` as its first instruction and executes `griddepcontrol.wait` before its first global
// memory access.  The next kernel's CTAs therefore become resident while the last wave of this one is still running, run
// their prologue (barrier init, TMEM allocation, descriptor prefetch) and start the moment this grid has completed and
// flushed -- the launch gap and the prologue disappear from the critical path; the data dependency is untouched because
// nothing is read or written before the wait.  Works in plain streams and under CUDA-graph capture (kernel -> kernel edges
// become programmatic edges).  STCGAN_PDL=0 turns the attribute off (the device instructions are then no-ops).
int pdl_enabled();

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  // a kernel launched on a high-priority stream carries that priority as a launch attribute, so that it survives CUDA-graph
  // capture (captured kernel nodes do not inherit the priority of the stream they were captured on)
  int prio = 0;
  if (cudaStreamGetPriority(st, &prio) == cudaSuccess && prio != 0) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = prio;
    ++na;
  }
  cfg.attrs = na ? attr : nullptr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through finish_launch()
}

#define STCGAN_REQUIRE(cond) do { if (!(cond)) return STCGAN_EINVAL; } while (0)

// ---------------------------------------------------------------------------------------------
// element types
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case STCGAN_ACT_LEAKY:   return z > 0.f ? z : 0.2f * z;
    case STCGAN_ACT_RELU:    return z > 0.f ? z : 0.f;
    case STCGAN_ACT_TANH:    return tanhf(z);
    case STCGAN_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default:                 return z;
  }
}
// Branch-free NONE / LeakyReLU(0.2) / ReLU for hot loops.  (act_fwd's switch also contains tanhf / expf; when the activation
// code is a run-time value nvcc if-converts the switch and evaluates EVERY branch per element -- measured: 90 cycles per
// element in the GEMM epilogue -- so streaming kernels must not call it.)
__device__ __forceinline__ float act_slope(int act) {
  return act == STCGAN_ACT_LEAKY ? 0.2f : (act == STCGAN_ACT_RELU ? 0.f : 1.f);
}
__device__ __forceinline__ float act_piecewise(float z, float slope) { return fmaxf(z, 0.f) + slope * fminf(z, 0.f); }

// derivative of LeakyReLU(0.2)/ReLU w.r.t. its input, evaluated from the pre-activation z
// (torch: leaky_relu_backward uses x > 0, threshold_backward uses x <= 0 -> 0)
__device__ __forceinline__ float act_gate(int act, float z) {
  return z > 0.f ? 1.f : act_slope(act);
}

// ---------------------------------------------------------------------------------------------
// tap geometry: out[n, a*os + oy0, b*os + ox0, :] = sum_j in[n, a*is + dy_j, b*is + dx_j, :] * Wp[wtap_j]
// ---------------------------------------------------------------------------------------------
struct Tap { int8_t dy, dx, wtap, view; };

struct Geom {
  int N, IH, IW, OH, OW;
  int istride, ostride;
  int nclass, ntaps;
  int8_t oy0[4], ox0[4];
  Tap tap[4][16];
  __host__ __device__ int grid_h(int c) const { return (OH - oy0[c] + ostride - 1) / ostride; }
  __host__ __device__ int grid_w(int c) const { return (OW - ox0[c] + ostride - 1) / ostride; }
};

// Build the geometry table for one of the STCGAN_GEOM_* kinds.  Returns false for an unknown kind.
inline bool make_geom(int kind, int N, int IH, int IW, int OH, int OW, Geom* g) {
  g->N = N; g->IH = IH; g->IW = IW; g->OH = OH; g->OW = OW;
  for (int c = 0; c < 4; ++c) { g->oy0[c] = 0; g->ox0[c] = 0; }
  switch (kind) {
    case STCGAN_GEOM_WIN_S2:
    case STCGAN_GEOM_WIN_S1:
    case STCGAN_GEOM_WIN_S1_FLIP: {
      const int s = kind == STCGAN_GEOM_WIN_S2 ? 2 : 1;
      const int pad = kind == STCGAN_GEOM_WIN_S1_FLIP ? 2 : 1;
      g->istride = s; g->ostride = 1; g->nclass = 1; g->ntaps = 16;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
          Tap& t = g->tap[0][a * 4 + b];
          t.dy = (int8_t)(a - pad); t.dx = (int8_t)(b - pad);
          t.wtap = (int8_t)(kind == STCGAN_GEOM_WIN_S1_FLIP ? (3 - a) * 4 + (3 - b) : a * 4 + b);
          // parity view of the input seen through TMA (stride-2 windows only): row 2a'+p
          const int ry = a - pad, rx = b - pad;
          t.view = (int8_t)(s == 2 ? ((ry & 1) * 2 + (rx & 1)) : 0);
        }
      return true;
    }
    case STCGAN_GEOM_PARITY: {
      // out row oy = 2a + ph receives kernel rows kh with (oy + 1 - kh) even, from input row (oy + 1 - kh)/2
      g->istride = 1; g->ostride = 2; g->nclass = 4; g->ntaps = 4;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const int c = ph * 2 + pw;
          g->oy0[c] = (int8_t)ph; g->ox0[c] = (int8_t)pw;
          int j = 0;
          for (int kh = 0; kh < 4; ++kh) {
            if (((ph + 1 - kh) & 1) != 0) continue;
            for (int kw = 0; kw < 4; ++kw) {
              if (((pw + 1 - kw) & 1) != 0) continue;
              Tap& t = g->tap[c][j++];
              // (2a + ph + 1 - kh)/2 = a + (ph + 1 - kh)/2   (exact: numerator even)
              t.dy = (int8_t)((ph + 1 - kh) / 2); t.dx = (int8_t)((pw + 1 - kw) / 2);
              t.wtap = (int8_t)(kh * 4 + kw); t.view = 0;
            }
          }
        }
      return true;
    }
    default: return false;
  }
}

// optional extras of a forward convolution's epilogue (tensor-core path):
//   v = acc * scale[c] + bias[c]  ->  y = act(v),  y2 = act2(v);  stores cropped to [HC, WC] (0 = the layer's output size)
struct EpilogueExtra {
  const float* scale = nullptr;      // per output channel (nullptr = 1); the per-channel shift travels as `bias`
  void* y2 = nullptr; int ldy2 = 0; int act2 = 0;
  int HC = 0, WC = 0;
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace stcgan

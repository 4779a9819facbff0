// BatchNorm2d statistics / apply / backward with the neighbouring LeakyReLU / ReLU fused in,
// on NHWC tensors.  All kernels are HBM-bound streaming passes with 128-bit accesses
// (8 bf16 or 4 fp32 channels per thread-load); statistics accumulate in fp64.
//
// Reference arithmetic: nn.BatchNorm2d (eps 1e-5, momentum 0.1, affine) + nn.LeakyReLU(0.2, True) /
// nn.ReLU(True) -- src/models/stcgan_g.py:87-90, src/models/stcgan_d.py:24,36-37,45-46.
#include <cstdlib>
#include "common.cuh"

namespace stcgan {

constexpr int VEC = 8;   // channels per thread (16 B of bf16, 2 x 16 B of fp32)

template <typename T> struct Vec8 { };
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// raw 8-channel loads, kept apart from the conversion so that several can be in flight per thread
template <typename T> struct Raw8 { };
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p + 4));
  }
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
  __device__ __forceinline__ void unpack(float* v) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void ld(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void zero() { u = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float* v) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};

// thread layout shared by all passes: CV = C/8 channel-vectors along x, pixels along y
struct RowMap {
  int cv, rows;   // threads per pixel row, pixel rows per block
};
static inline RowMap row_map(int C) {
  RowMap m; m.cv = C / VEC; if (m.cv > 256) m.cv = 256;
  m.rows = 256 / m.cv; if (m.rows < 1) m.rows = 1;
  return m;
}

// ---------------------------------------------------------------------------------------------
// statistics: acc[0][c] += sum y, acc[1][c] += sum y^2     (also used for bias grads via colsum)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ y, long long P, int C, int ld, double* __restrict__ acc,
                int cv, int rows, int want_sq) {
  pdl_prologue();
  extern __shared__ float red[];   // [rows][cv*8] x2
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  const bool active = tr < rows;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    // one channel-vector column per outer iteration (C > 2048 only)
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
    if (active) {
      const long long stride = (long long)gridDim.x * rows;
      for (long long p = (long long)blockIdx.x * rows + tr; p < P; p += 4 * stride) {
        Raw8<T> raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {          // unconditional (clamped) loads, all issued before any is consumed
          const long long pp = p + u * stride;
          raw[u].ld(y + (pp < P ? pp : P - 1) * ld + c0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (p + u * stride < P) {
            float v[8];
            raw[u].unpack(v);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
          }
      }
    }
    float* rs = red; float* rq = red + rows * cv * VEC;
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { rs[(tr * cv + tc) * VEC + i] = s[i]; rq[(tr * cv + tc) * VEC + i] = q[i]; }
    }
    __syncthreads();
    // first cv*8 threads (or loop) reduce over rows in double
    for (int j = threadIdx.x; j < cv * VEC; j += blockDim.x) {
      double ds = 0.0, dq = 0.0;
      for (int r = 0; r < rows; ++r) { ds += rs[r * cv * VEC + j]; dq += rq[r * cv * VEC + j]; }
      const int c = (c0 - tc * VEC) + j;
      atomicAdd(&acc[c], ds);
      if (want_sq) atomicAdd(&acc[C + c], dq);
    }
    __syncthreads();
  }
}

// column sums for any channel count (bias gradients: C = 1, 3 or 64): channels along x, pixels along y
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ g, long long P, int C, int ld, float* __restrict__ out, int cc, int rows) {
  pdl_prologue();
  __shared__ float red[256];
  const int tc = threadIdx.x % cc, tr = threadIdx.x / cc;
  for (int c0 = 0; c0 < C; c0 += cc) {
    const int c = c0 + tc;
    float s = 0.f;
    if (tr < rows && c < C)
      for (long long p = (long long)blockIdx.x * rows + tr; p < P; p += (long long)gridDim.x * rows)
        s += to_f32<T>(g[p * ld + c]);
    red[threadIdx.x] = s;
    __syncthreads();
    if (tr == 0 && c < C) {
      float t = 0.f;
      for (int r = 0; r < rows; ++r) t += red[r * cc + tc];
      atomicAdd(&out[c], t);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ acc, long long P, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                   int training, float* __restrict__ mean_invstd, float* __restrict__ scale_shift) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    const double m = acc[c] / (double)P;
    double var = acc[C + c] / (double)P - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) {
      const double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
  } else {
    mean = rmean[c];
    invstd = 1.f / sqrtf(rvar[c] + eps);
  }
  mean_invstd[c] = mean; mean_invstd[C + c] = invstd;
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc; scale_shift[C + c] = beta[c] - mean * sc;
}

// deferred running-statistics update: the momentum update that bn_fused_apply_kernel's block 0 performs, as a launch of
// its own (same arithmetic, bit for bit: sums over the statistic slots in fp64, 1/count and count/(count-1) from the
// host).  Lets two training forward passes of one network run concurrently -- the second one skips its update and
// applies it afterwards, which keeps the reference's update ORDER (the exponential average does not commute).
__global__ void bn_running_update_kernel(const double* __restrict__ acc, int C, double inv_count, double unbias,
                                         float* __restrict__ rmean, float* __restrict__ rvar, float momentum) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
#pragma unroll
  for (int k = 0; k < STCGAN_BN_SLOTS; ++k) { s += acc[(2 * k) * C + c]; q += acc[(2 * k + 1) * C + c]; }
  const double m = s * inv_count;
  double var = q * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  const double unbiased = var * unbias;
  rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
  rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
}

int bn_running_update(const double* acc, long long count, int C, float* rmean, float* rvar, float momentum, cudaStream_t st) {
  const double inv_count = count > 0 ? 1.0 / (double)count : 0.0;
  const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  launch_k(bn_running_update_kernel, (C + 127) / 128, 128, 0, st, acc, C, inv_count, unbias, rmean, rvar, momentum);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------------
// forward apply (+ activation, optional second output with another activation, optional crop)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const T* __restrict__ y, int H, int W, int C, int ldy, const float* __restrict__ ss,
                    int HC, int WC, long long PC, T* __restrict__ o1, int ld1, int act1,
                    T* __restrict__ o2, int ld2, int act2, int cv, int rows) {
  pdl_prologue();
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  if (tr >= rows) return;
  const float slope1 = act_slope(act1), slope2 = act_slope(act2);
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ss ? ss[c0 + i] : 1.f; sh[i] = ss ? ss[C + c0 + i] : 0.f; }
    const long long stride = (long long)gridDim.x * rows;
    const bool nocrop = HC == H && WC == W;
    for (long long p0 = (long long)blockIdx.x * rows + tr; p0 < PC; p0 += 4 * stride) {
      Raw8<T> raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        long long p = p0 + u * stride;
        if (p >= PC) p = PC - 1;                    // clamped: the load is unconditional, the store is not
        long long src = p;
        if (!nocrop) {
          const int w = (int)(p % WC); const long long t = p / WC;
          const int h = (int)(t % HC); const long long n = t / HC;
          src = (n * H + h) * W + w;
        }
        raw[u].ld(y + src * (long long)ldy + c0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long p = p0 + u * stride;
        if (p < PC) {
          Vec8<T> a, b;
          float yv[8];
          raw[u].unpack(yv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = fmaf(yv[i], sc[i], sh[i]);
            a.v[i] = act_piecewise(z, slope1);
            b.v[i] = act_piecewise(z, slope2);
          }
          a.store(o1 + p * ld1 + c0);
          if (o2) b.store(o2 + p * ld2 + c0);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// finalize + apply in one launch: every block derives scale / shift of all channels from the statistics (the sums of
// STCGAN_BN_SLOTS partial slots, training) or from the running statistics (eval) into shared memory; block 0 also
// publishes mean / invstd / scale / shift for the backward pass and updates the running statistics.
// ---------------------------------------------------------------------------------------------
struct BnFinalize {
  const double* acc;      // [STCGAN_BN_SLOTS][2][C] (training)
  long long count;        // values per channel
  double inv_count, unbias;   // 1 / count, count / (count - 1)
  const float* gamma; const float* beta;
  float* rmean; float* rvar;
  float momentum, eps;
  int training;
  float* mean_invstd; float* scale_shift;   // [2][C] each
};

template <typename T, int U = 4, int MINB = 1>
__global__ void __launch_bounds__(256, MINB)
bn_fused_apply_kernel(const T* __restrict__ y, int H, int W, int C, int ldy, const BnFinalize f,
                      int HC, int WC, long long PC, T* __restrict__ o1, int ld1, int act1,
                      T* __restrict__ o2, int ld2, int act2, int cv, int rows) {
  extern __shared__ float ssm[];     // [2][C]: scale, shift
  pdl_prologue();
  const double inv_count = f.inv_count, unbias = f.unbias;      // formed on the host
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    if (f.training) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int k = 0; k < STCGAN_BN_SLOTS; ++k) { s += f.acc[(2 * k) * C + c]; q += f.acc[(2 * k + 1) * C + c]; }
      // (fp64 only where cancellation needs it: sums, mean, E[y^2] - mean^2; B200's fp64 divide / sqrt are emulated and
      // cost microseconds in these latency-bound launches, so 1/count is formed once per thread and invstd in fp32, which
      // is also what torch's float kernels compute)
      const double m = s * inv_count;
      double var = q * inv_count - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = 1.f / sqrtf((float)(var + (double)f.eps));
      if (blockIdx.x == 0 && f.rmean) {
        const double unbiased = var * unbias;
        f.rmean[c] = (1.f - f.momentum) * f.rmean[c] + f.momentum * (float)m;
        f.rvar[c] = (1.f - f.momentum) * f.rvar[c] + f.momentum * (float)unbiased;
      }
    } else {
      mean = f.rmean[c];
      invstd = 1.f / sqrtf(f.rvar[c] + f.eps);
    }
    const float sc = f.gamma[c] * invstd, sh = f.beta[c] - mean * sc;
    ssm[c] = sc; ssm[C + c] = sh;
    if (blockIdx.x == 0) {
      f.mean_invstd[c] = mean; f.mean_invstd[C + c] = invstd;
      f.scale_shift[c] = sc; f.scale_shift[C + c] = sh;
    }
  }
  __syncthreads();
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  if (tr >= rows) return;
  const float slope1 = act_slope(act1), slope2 = act_slope(act2);
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ssm[c0 + i]; sh[i] = ssm[C + c0 + i]; }
    const long long stride = (long long)gridDim.x * rows;
    const bool nocrop = HC == H && WC == W;
    for (long long p0 = (long long)blockIdx.x * rows + tr; p0 < PC; p0 += U * stride) {
      Raw8<T> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        long long p = p0 + u * stride;
        if (p >= PC) p = PC - 1;
        long long src = p;
        if (!nocrop) {
          const int w = (int)(p % WC); const long long t = p / WC;
          const int h = (int)(t % HC); const long long n = t / HC;
          src = (n * H + h) * W + w;
        }
        raw[u].ld(y + src * (long long)ldy + c0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * stride;
        if (p < PC) {
          Vec8<T> a, b;
          float yv[8];
          raw[u].unpack(yv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = fmaf(yv[i], sc[i], sh[i]);
            a.v[i] = act_piecewise(z, slope1);
            b.v[i] = act_piecewise(z, slope2);
          }
          a.store(o1 + p * ld1 + c0);
          if (o2) b.store(o2 + p * ld2 + c0);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// raw operands of one pixel (8 channels): y, and the incoming gradients (zero outside the crop).  `load` only issues
// the (unconditional, clamped) global loads; `dz` converts and combines, so loads of several pixels overlap.
template <typename T>
struct BwdIn {
  Raw8<T> ry, ra, rb;
  bool inside;
  __device__ __forceinline__ void load(const T* yp, long long P, int H, int W, int ldy, int HC, int WC, const T* g1, int ldg1,
                                       const T* g2, int ldg2, long long p, int c0) {
    if (p >= P) p = P - 1;
    ry.ld(yp + p * ldy + c0);
    long long pc = p;
    inside = true;
    if (HC != H || WC != W) {            // p indexes the FULL [N,H,W] grid; gradients live on the cropped grid
      const int w = (int)(p % W); const long long t = p / W;
      const int h = (int)(t % H); const long long n = t / H;
      inside = h < HC && w < WC;
      pc = (n * HC + (h < HC ? h : HC - 1)) * WC + (w < WC ? w : WC - 1);
    }
    ra.ld(g1 + pc * ldg1 + c0);
    if (g2) rb.ld(g2 + pc * ldg2 + c0); else rb.zero();
  }
  // dz = g1*act1'(z) + g2*act2'(z),  z = y*sc + sh ; also returns y
  __device__ __forceinline__ void dz(const float* sc, const float* sh, int act1, int act2, bool two, float* out, float* yv) const {
    float av[8], bv[8];
    ry.unpack(yv); ra.unpack(av); rb.unpack(bv);
    const float s1 = act_slope(act1), s2 = act_slope(act2);
    const float m = inside ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float z = fmaf(yv[i], sc[i], sh[i]);
      out[i] = m * (av[i] * (z > 0.f ? 1.f : s1) + (two ? bv[i] * (z > 0.f ? 1.f : s2) : 0.f));
    }
  }
};

// pass 1: acc[slot][0][c] += sum dz ; acc[slot][1][c] += sum dz * (y - mean)   (fp64 across threads / blocks; the blocks
// spread their atomics over STCGAN_BN_SLOTS partial slots, pass 2 adds the slots up)
template <typename T, int U = 4, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_reduce_kernel(const T* __restrict__ y, long long P, int H, int W, int C, int ldy,
                     const float* __restrict__ ss, const float* __restrict__ mi, int HC, int WC,
                     const T* __restrict__ g1, int ldg1, int act1, const T* __restrict__ g2, int ldg2, int act2,
                     double* __restrict__ acc, int cv, int rows) {
  extern __shared__ float red[];
  pdl_prologue();
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  const bool active = tr < rows;
  const bool two = g2 != nullptr;
  double* acc_slot = acc + (long long)(blockIdx.x % STCGAN_BN_SLOTS) * 2 * C;
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8], mean[8], s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ss[c0 + i]; sh[i] = ss[C + c0 + i]; mean[i] = mi[c0 + i]; s[i] = 0.f; q[i] = 0.f; }
    if (active) {
      const long long stride = (long long)gridDim.x * rows;
      for (long long p = (long long)blockIdx.x * rows + tr; p < P; p += U * stride) {
        BwdIn<T> in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) in[u].load(y, P, H, W, ldy, HC, WC, g1, ldg1, g2, ldg2, p + u * stride, c0);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (p + u * stride < P) {
            float dz[8], yv[8];
            in[u].dz(sc, sh, act1, act2, two, dz, yv);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += dz[i]; q[i] = fmaf(dz[i], yv[i] - mean[i], q[i]); }
          }
      }
    }
    float* rs = red; float* rq = red + rows * cv * VEC;
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { rs[(tr * cv + tc) * VEC + i] = s[i]; rq[(tr * cv + tc) * VEC + i] = q[i]; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cv * VEC; j += blockDim.x) {
      double ds = 0.0, dq = 0.0;
      for (int r = 0; r < rows; ++r) { ds += rs[r * cv * VEC + j]; dq += rq[r * cv * VEC + j]; }
      const int c = (c0 - tc * VEC) + j;
      atomicAdd(&acc_slot[c], ds);
      atomicAdd(&acc_slot[C + c], dq);
    }
    __syncthreads();
  }
}

// pass 2: dy = A*dz - B - (y - mean)*Cc  with A = gamma*invstd, B = A*mean(dz), Cc = A*invstd^2*mean(dz*(y-mean))
//         (training);  dy = A*dz (eval);  dy = dz (no BatchNorm).  dgamma += invstd*acc[1], dbeta += acc[0] (block 0,
//         atomically: the real and the fake pass of a discriminator may run their backward passes concurrently).
//         The per-channel coefficients are derived once per block into shared memory (sum over the statistic slots).
//         dbias (optional, layers with a conv bias and no BatchNorm): dbias[c] += sum_p dy[p, c].
template <typename T, int U = 4, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_apply_kernel(const T* __restrict__ y, long long P, int H, int W, int C, int ldy,
                    const float* __restrict__ ss, const float* __restrict__ mi, const float* __restrict__ gamma,
                    int training, int HC, int WC,
                    const T* __restrict__ g1, int ldg1, int act1, const T* __restrict__ g2, int ldg2, int act2,
                    const double* __restrict__ acc, T* __restrict__ dy, int lddy,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int cv, int rows,
                    double invP) {
  extern __shared__ float coef[];     // [5][C]: sc, sh, kA, kB, kC   (+ [rows][cv*8] bias-gradient scratch behind it)
  pdl_prologue();
  const bool has_bn = ss != nullptr;
  const bool two = g2 != nullptr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sc = 1.f, sh = 0.f, kA = 1.f, kB = 0.f, kC = 0.f;
    if (has_bn) {
      sc = ss[c]; sh = ss[C + c];
      const float mean = mi[c], invstd = mi[C + c];
      kA = gamma[c] * invstd;
      if (training) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < STCGAN_BN_SLOTS; ++k) { a0 += acc[(2 * k) * C + c]; a1 += acc[(2 * k + 1) * C + c]; }
        kC = kA * invstd * invstd * (float)(a1 * invP);
        kB = kA * (float)(a0 * invP) - mean * kC;        // folded: -(y - mean)*kC = -y*kC + mean*kC
        if (blockIdx.x == 0 && dgamma) { atomicAdd(&dgamma[c], (float)(a1 * (double)invstd)); atomicAdd(&dbeta[c], (float)a0); }
      } else if (blockIdx.x == 0 && dgamma && acc) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < STCGAN_BN_SLOTS; ++k) { a0 += acc[(2 * k) * C + c]; a1 += acc[(2 * k + 1) * C + c]; }
        atomicAdd(&dgamma[c], (float)(a1 * (double)invstd)); atomicAdd(&dbeta[c], (float)a0);
      }
    }
    coef[c] = sc; coef[C + c] = sh; coef[2 * C + c] = kA; coef[3 * C + c] = kB; coef[4 * C + c] = kC;
  }
  __syncthreads();
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  const bool active = tr < rows;
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8], kA[8], kB[8], kC[8], bs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = coef[c0 + i]; sh[i] = coef[C + c0 + i]; kA[i] = coef[2 * C + c0 + i]; kB[i] = coef[3 * C + c0 + i];
      kC[i] = coef[4 * C + c0 + i]; bs[i] = 0.f;
    }
    if (active) {
      const long long stride = (long long)gridDim.x * rows;
      for (long long p = (long long)blockIdx.x * rows + tr; p < P; p += U * stride) {
        BwdIn<T> in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) in[u].load(y, P, H, W, ldy, HC, WC, g1, ldg1, g2, ldg2, p + u * stride, c0);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (p + u * stride < P) {
            float dz[8], yv[8];
            in[u].dz(sc, sh, act1, act2, two, dz, yv);
            Vec8<T> o;
#pragma unroll
            for (int i = 0; i < 8; ++i) { o.v[i] = fmaf(kA[i], dz[i], -kB[i]) - yv[i] * kC[i]; bs[i] += o.v[i]; }
            o.store(dy + (p + u * stride) * lddy + c0);
          }
      }
    }
    if (dbias) {      // block-level column sums, one fp32 atomic per channel and block
      float* rb = coef + 5 * C;
      if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[(tr * cv + tc) * VEC + i] = bs[i];
      }
      __syncthreads();
      for (int j = threadIdx.x; j < cv * VEC; j += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += rb[r * cv * VEC + j];
        atomicAdd(&dbias[(c0 - tc * VEC) + j], t);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tiny tensors (the innermost U-Net levels: <= 64 pixels x 512 channels): reduce + apply in ONE single-block launch.  The
// two-pass form costs two dependent ~7 us launches -- pure latency on the backward pass's critical chain; one 512-thread
// block (<= 128 registers per thread: at 1024 threads the 64-register cap made the pixel loop spill) reduces in shared
// memory, derives the coefficients and applies them (the second read of y / g hits L1/L2).  One SM streams ~100 GB/s, so
// this only wins below ~0.5 MB of traffic: measured on B200, 256-pixel tensors (1.8 MB) took 30 us this way against 14 us
// in two passes, hence the 64-pixel bound.  Same arithmetic as the two-pass kernels: fp32 partial sums per thread, fp64
// across the block.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512, 1)
bn_bwd_small_kernel(const T* __restrict__ y, long long P, int H, int W, int C, int ldy,
                    const float* __restrict__ ss, const float* __restrict__ mi, const float* __restrict__ gamma,
                    int HC, int WC, const T* __restrict__ g1, int ldg1, int act1, const T* __restrict__ g2, int ldg2, int act2,
                    T* __restrict__ dy, int lddy, float* __restrict__ dgamma, float* __restrict__ dbeta, int cv, int rows,
                    double invP) {
  extern __shared__ float sm_small[];       // [2][rows][C] partial sums, then [3][C] coefficients behind them
  pdl_prologue();
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv;
  const bool active = tr < rows;
  const bool two = g2 != nullptr;
  float* rs = sm_small; float* rq = sm_small + rows * C;
  float* coef = sm_small + 2 * rows * C;    // kA, kB, kC
  // ---- pass 1: per-thread sums over this thread's pixels, every channel vector this thread owns (C <= cv * 8 * k)
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8], mean[8], s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ss[c0 + i]; sh[i] = ss[C + c0 + i]; mean[i] = mi[c0 + i]; s[i] = 0.f; q[i] = 0.f; }
    if (active)
      for (long long p = tr; p < P; p += 2 * rows) {
        BwdIn<T> in[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) in[u].load(y, P, H, W, ldy, HC, WC, g1, ldg1, g2, ldg2, p + (long long)u * rows, c0);
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (p + (long long)u * rows < P) {
            float dz[8], yv[8];
            in[u].dz(sc, sh, act1, act2, two, dz, yv);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += dz[i]; q[i] = fmaf(dz[i], yv[i] - mean[i], q[i]); }
          }
      }
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { rs[tr * C + c0 + i] = s[i]; rq[tr * C + c0 + i] = q[i]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < rows; ++r) { a0 += rs[r * C + c]; a1 += rq[r * C + c]; }
    const float mean = mi[c], invstd = mi[C + c];
    const float kA = gamma[c] * invstd;
    const float kC = kA * invstd * invstd * (float)(a1 * invP);
    coef[c] = kA; coef[C + c] = kA * (float)(a0 * invP) - mean * kC; coef[2 * C + c] = kC;
    if (dgamma) { atomicAdd(&dgamma[c], (float)(a1 * (double)invstd)); atomicAdd(&dbeta[c], (float)a0); }
  }
  __syncthreads();
  // ---- pass 2: dy = kA*dz - kB - y*kC
  for (int c0 = tc * VEC; c0 < C; c0 += cv * VEC) {
    float sc[8], sh[8], kA[8], kB[8], kC[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = ss[c0 + i]; sh[i] = ss[C + c0 + i]; kA[i] = coef[c0 + i]; kB[i] = coef[C + c0 + i]; kC[i] = coef[2 * C + c0 + i];
    }
    if (active)
      for (long long p = tr; p < P; p += 2 * rows) {
        BwdIn<T> in[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) in[u].load(y, P, H, W, ldy, HC, WC, g1, ldg1, g2, ldg2, p + (long long)u * rows, c0);
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (p + (long long)u * rows < P) {
            float dz[8], yv[8];
            in[u].dz(sc, sh, act1, act2, two, dz, yv);
            Vec8<T> o;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.v[i] = fmaf(kA[i], dz[i], -kB[i]) - yv[i] * kC[i];
            o.store(dy + (p + (long long)u * rows) * lddy + c0);
          }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Grid for a grid-stride streaming kernel: exactly one wave of resident CTAs (148 SMs x occupancy), fewer for small
// inputs.  `max_per_sm` caps it further for kernels that end in per-block atomics (fewer, fatter blocks).
template <typename K>
static inline unsigned stream_grid(long long P, int rows, K kernel, size_t smem, int max_per_sm = 8) {
  // occupancy is queried once per kernel (before any CUDA-graph capture, during the warm-up steps) and cached
  static const void* keys[32]; static int vals[32]; static int nkeys = 0;
  int occ = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < nkeys; ++i) if (keys[i] == key) occ = vals[i];
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 1;
    if (nkeys < 32) { keys[nkeys] = key; vals[nkeys] = occ; ++nkeys; }
  }
  if (occ > max_per_sm) occ = max_per_sm;
  long long b = (P + rows - 1) / rows;
  const long long cap = 148LL * occ;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// (Measured on B200, round 2: 4 pixels in flight per thread at 1-2 blocks/SM -- the shapes below -- beat 2 in flight at 3-4
// blocks/SM and 4 in flight at 3 blocks/SM by 5-20 % of the whole train step: the register-starved variants spill.)
template <typename T>
static int bn_stats_t(const void* y, long long P, int C, int ld, double* acc, int want_sq, cudaStream_t st) {
  const RowMap m = row_map(C);
  const size_t smem = (size_t)2 * m.rows * m.cv * VEC * sizeof(float);
  launch_k(bn_stats_kernel<T>, stream_grid(P, m.rows * 8, bn_stats_kernel<T>, smem, 2), 256, smem, st, static_cast<const T*>(y), P, C, ld, acc, m.cv, m.rows, want_sq);
  return finish_launch();
}

int bn_stats(int dtype, const void* y, long long P, int C, int ld, double* acc, cudaStream_t st) {
  const int esz = dtype == STCGAN_F32 ? 4 : 2;
  if (C % VEC != 0 || C > 2048 || !aligned16(y) || ((long long)ld * esz) % 16 != 0) return STCGAN_EINVAL;
  if (P == 0) return 0;
  return dtype == STCGAN_F32 ? bn_stats_t<float>(y, P, C, ld, acc, 1, st) : bn_stats_t<__nv_bfloat16>(y, P, C, ld, acc, 1, st);
}

int bn_finalize(const double* acc, long long P, int C, const float* gamma, const float* beta, float* rmean, float* rvar,
                float momentum, float eps, int training, float* mean_invstd, float* scale_shift, cudaStream_t st) {
  if (!training && (!rmean || !rvar)) return STCGAN_EINVAL;
  launch_k(bn_finalize_kernel, (C + 127) / 128, 128, 0, st, acc, P, C, gamma, beta, rmean, rvar, momentum, eps, training,
                                                      mean_invstd, scale_shift);
  return finish_launch();
}

static bool vec_ok(int dtype, const void* p, int ld) {
  const int esz = dtype == STCGAN_F32 ? 4 : 2;
  return p == nullptr || (aligned16(p) && ((long long)ld * esz) % 16 == 0);
}

int bn_act_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, int HC, int WC,
                 void* o1, int ld1, int act1, void* o2, int ld2, int act2, cudaStream_t st) {
  if (C % VEC != 0 || C > 2048 || HC > H || WC > W || !vec_ok(dtype, y, ldy) || !vec_ok(dtype, o1, ld1) || !vec_ok(dtype, o2, ld2) || !o1)
    return STCGAN_EINVAL;
  const long long PC = (long long)N * HC * WC;
  if (PC == 0) return 0;
  const RowMap m = row_map(C);
  if (dtype == STCGAN_F32)
    launch_k(bn_act_apply_kernel<float>, stream_grid(PC, m.rows * 4, bn_act_apply_kernel<float>, 0), 256, 0, st, static_cast<const float*>(y), H, W, C, ldy, ss, HC, WC, PC, static_cast<float*>(o1), ld1, act1,
        static_cast<float*>(o2), ld2, act2, m.cv, m.rows);
  else
    launch_k(bn_act_apply_kernel<__nv_bfloat16>, stream_grid(PC, m.rows * 4, bn_act_apply_kernel<__nv_bfloat16>, 0), 256, 0, st, static_cast<const __nv_bfloat16*>(y), H, W, C, ldy, ss, HC, WC, PC, static_cast<__nv_bfloat16*>(o1), ld1, act1,
        static_cast<__nv_bfloat16*>(o2), ld2, act2, m.cv, m.rows);
  return finish_launch();
}

int bn_fused_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const double* acc, long long count,
                   const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps, int training,
                   float* mean_invstd, float* scale_shift, int HC, int WC,
                   void* o1, int ld1, int act1, void* o2, int ld2, int act2, cudaStream_t st) {
  if (C % VEC != 0 || C > 2048 || HC > H || WC > W || !vec_ok(dtype, y, ldy) || !vec_ok(dtype, o1, ld1) || !vec_ok(dtype, o2, ld2) || !o1)
    return STCGAN_EINVAL;
  if (!gamma || !beta || !mean_invstd || !scale_shift) return STCGAN_EINVAL;
  if (training ? (!acc || count < 1) : (!rmean || !rvar)) return STCGAN_EINVAL;
  const long long PC = (long long)N * HC * WC;
  if (PC == 0) return 0;
  const RowMap m = row_map(C);
  BnFinalize f;
  f.inv_count = count > 0 ? 1.0 / (double)count : 0.0;
  f.unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  f.acc = acc; f.count = count; f.gamma = gamma; f.beta = beta; f.rmean = rmean; f.rvar = rvar;
  f.momentum = momentum; f.eps = eps; f.training = training; f.mean_invstd = mean_invstd; f.scale_shift = scale_shift;
  const size_t smem = (size_t)2 * C * sizeof(float);
  if (dtype == STCGAN_F32)
    launch_k(bn_fused_apply_kernel<float>, stream_grid(PC, m.rows * 4, bn_fused_apply_kernel<float>, smem), 256, smem, st,
             static_cast<const float*>(y), H, W, C, ldy, f, HC, WC, PC, static_cast<float*>(o1), ld1, act1,
             static_cast<float*>(o2), ld2, act2, m.cv, m.rows);
  else
    launch_k(bn_fused_apply_kernel<__nv_bfloat16>, stream_grid(PC, m.rows * 4, bn_fused_apply_kernel<__nv_bfloat16>, smem), 256, smem, st,
             static_cast<const __nv_bfloat16*>(y), H, W, C, ldy, f, HC, WC, PC, static_cast<__nv_bfloat16*>(o1), ld1, act1,
             static_cast<__nv_bfloat16*>(o2), ld2, act2, m.cv, m.rows);
  return finish_launch();
}

int bn_act_bwd_reduce(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                      int HC, int WC, const void* g1, int ldg1, int act1, const void* g2, int ldg2, int act2,
                      double* acc, cudaStream_t st) {
  if (C % VEC != 0 || C > 2048 || !ss || !mi || !g1 || !vec_ok(dtype, y, ldy) || !vec_ok(dtype, g1, ldg1) || !vec_ok(dtype, g2, ldg2))
    return STCGAN_EINVAL;
  const long long P = (long long)N * H * W;
  if (P == 0) return 0;
  const RowMap m = row_map(C);
  const size_t smem = (size_t)2 * m.rows * m.cv * VEC * sizeof(float);
  if (dtype == STCGAN_F32)
    launch_k(bn_bwd_reduce_kernel<float>, stream_grid(P, m.rows * 4, bn_bwd_reduce_kernel<float>, smem, 4), 256, smem, st, static_cast<const float*>(y), P, H, W, C, ldy, ss, mi, HC, WC, static_cast<const float*>(g1), ldg1, act1,
        static_cast<const float*>(g2), ldg2, act2, acc, m.cv, m.rows);
  else
    launch_k(bn_bwd_reduce_kernel<__nv_bfloat16>, stream_grid(P, m.rows * 4, bn_bwd_reduce_kernel<__nv_bfloat16>, smem, 4), 256, smem, st, static_cast<const __nv_bfloat16*>(y), P, H, W, C, ldy, ss, mi, HC, WC, static_cast<const __nv_bfloat16*>(g1), ldg1,
        act1, static_cast<const __nv_bfloat16*>(g2), ldg2, act2, acc, m.cv, m.rows);
  return finish_launch();
}

int bn_act_bwd_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                     const float* gamma, int training, int HC, int WC, const void* g1, int ldg1, int act1,
                     const void* g2, int ldg2, int act2, const double* acc, void* dy, int lddy,
                     float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  if (C % VEC != 0 || C > 2048 || !g1 || !dy || !vec_ok(dtype, y, ldy) || !vec_ok(dtype, g1, ldg1) || !vec_ok(dtype, g2, ldg2) ||
      !vec_ok(dtype, dy, lddy))
    return STCGAN_EINVAL;
  if (ss && (!mi || !gamma || (training && !acc))) return STCGAN_EINVAL;
  const long long P = (long long)N * H * W;
  if (P == 0) return 0;
  const RowMap m = row_map(C);
  const size_t smem = (size_t)5 * C * sizeof(float) + (dbias ? (size_t)m.rows * m.cv * VEC * sizeof(float) : 0);
  if (dtype == STCGAN_F32)
    launch_k(bn_bwd_apply_kernel<float>, stream_grid(P, m.rows * 4, bn_bwd_apply_kernel<float>, smem), 256, smem, st,
             static_cast<const float*>(y), P, H, W, C, ldy, ss, mi, gamma, training, HC, WC, static_cast<const float*>(g1), ldg1,
             act1, static_cast<const float*>(g2), ldg2, act2, acc, static_cast<float*>(dy), lddy, dgamma, dbeta, dbias, m.cv, m.rows,
             1.0 / (double)P);
  else
    launch_k(bn_bwd_apply_kernel<__nv_bfloat16>, stream_grid(P, m.rows * 4, bn_bwd_apply_kernel<__nv_bfloat16>, smem), 256, smem, st,
             static_cast<const __nv_bfloat16*>(y), P, H, W, C, ldy, ss, mi, gamma, training, HC, WC,
             static_cast<const __nv_bfloat16*>(g1), ldg1, act1, static_cast<const __nv_bfloat16*>(g2), ldg2, act2, acc,
             static_cast<__nv_bfloat16*>(dy), lddy, dgamma, dbeta, dbias, m.cv, m.rows, 1.0 / (double)P);
  return finish_launch();
}

// training-mode BatchNorm(+activation) backward of a SMALL tensor in one launch; returns STCGAN_EUNSUPPORTED when the tensor
// is not small (the caller then runs the two-pass form)
int bn_act_bwd_small(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                     const float* gamma, int HC, int WC, const void* g1, int ldg1, int act1, const void* g2, int ldg2,
                     int act2, void* dy, int lddy, float* dgamma, float* dbeta, cudaStream_t st) {
  const long long P = (long long)N * H * W;
  if (dtype != STCGAN_BF16 || C % VEC != 0 || C > 4096 || P * C > 64LL * 512 || P < 1) return STCGAN_EUNSUPPORTED;
  if (!ss || !mi || !gamma || !g1 || !dy || !vec_ok(dtype, y, ldy) || !vec_ok(dtype, g1, ldg1) || !vec_ok(dtype, g2, ldg2) ||
      !vec_ok(dtype, dy, lddy))
    return STCGAN_EINVAL;
  int cv = C / VEC; if (cv > 512) cv = 512;
  int rows = 512 / cv; if (rows < 1) rows = 1;
  if ((long long)rows > P) rows = (int)P;
  const size_t smem = ((size_t)2 * rows * C + 3 * C) * sizeof(float);
  if (smem > 200 * 1024) return STCGAN_EUNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(bn_bwd_small_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(bn_bwd_small_kernel<__nv_bfloat16>, 1, 512, smem, st, static_cast<const __nv_bfloat16*>(y), P, H, W, C, ldy, ss, mi,
           gamma, HC, WC, static_cast<const __nv_bfloat16*>(g1), ldg1, act1, static_cast<const __nv_bfloat16*>(g2), ldg2, act2,
           static_cast<__nv_bfloat16*>(dy), lddy, dgamma, dbeta, cv, rows, 1.0 / (double)P);
  return finish_launch();
}

int colsum(int dtype, const void* g, long long P, int C, int ld, float* out, cudaStream_t st) {
  if (P == 0 || C == 0) return 0;
  const int cc = C < 256 ? C : 256, rows = 256 / cc;
  long long b = (P + rows * 8 - 1) / (rows * 8); if (b > 148 * 8) b = 148 * 8; if (b < 1) b = 1;
  if (dtype == STCGAN_F32)
    launch_k(colsum_kernel<float>, (unsigned)b, 256, 0, st, static_cast<const float*>(g), P, C, ld, out, cc, rows);
  else
    launch_k(colsum_kernel<__nv_bfloat16>, (unsigned)b, 256, 0, st, static_cast<const __nv_bfloat16*>(g), P, C, ld, out, cc, rows);
  return finish_launch();
}

}  // namespace stcgan

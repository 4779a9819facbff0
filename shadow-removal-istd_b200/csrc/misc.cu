// Layout, loss, optimiser and quantisation kernels (all HBM-/latency-bound streaming work).
#include "common.cuh"

namespace stcgan {

// ---------------------------------------------------------------------------------------------
// weight pack: W[d0][d1][16] fp32 -> P1[t][d0][d1], P2[t][d1][d0]  (T)
// one block = 16 d0 x 32 d1 tile, all 16 taps; smem transposes so that every global access is coalesced
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, int D0, int D1, T* __restrict__ p1, T* __restrict__ p2) {
  pdl_prologue();
  __shared__ float tile[16][32 * 16 + 1];   // [d0][d1*16 + t]
  const int a0 = blockIdx.y * 16, b0 = blockIdx.x * 32;
  // load: rows of 32*16 contiguous floats per d0
  for (int i = threadIdx.x; i < 16 * 512; i += 256) {
    const int r = i / 512, c = i % 512;
    const int d0 = a0 + r, d1 = b0 + c / 16;
    tile[r][c] = (d0 < D0 && d1 < D1) ? w[((long long)d0 * D1 + b0) * 16 + c] : 0.f;
  }
  __syncthreads();
  if (p1) {
    for (int i = threadIdx.x; i < 16 * 16 * 32; i += 256) {
      const int t = i / 512, r = (i / 32) % 16, c = i % 32;   // consecutive threads -> consecutive d1
      const int d0 = a0 + r, d1 = b0 + c;
      if (d0 < D0 && d1 < D1) p1[((long long)t * D0 + d0) * D1 + d1] = from_f32<T>(tile[r][c * 16 + t]);
    }
  }
  if (p2) {
    for (int i = threadIdx.x; i < 16 * 16 * 32; i += 256) {
      const int t = i / 512, c = (i / 16) % 32, r = i % 16;   // consecutive threads -> consecutive d0
      const int d0 = a0 + r, d1 = b0 + c;
      if (d0 < D0 && d1 < D1) p2[((long long)t * D1 + d1) * D0 + d0] = from_f32<T>(tile[r][c * 16 + t]);
    }
  }
}

int pack_weight(int dtype, const float* w, int D0, int D1, void* p1, void* p2, cudaStream_t st) {
  if (D0 <= 0 || D1 <= 0) return STCGAN_EINVAL;
  dim3 grid((D1 + 31) / 32, (D0 + 15) / 16);
  if (dtype == STCGAN_F32)
    launch_k(pack_weight_kernel<float>, grid, 256, 0, st, w, D0, D1, static_cast<float*>(p1), static_cast<float*>(p2));
  else
    launch_k(pack_weight_kernel<__nv_bfloat16>, grid, 256, 0, st, w, D0, D1, static_cast<__nv_bfloat16*>(p1),
                                                            static_cast<__nv_bfloat16*>(p2));
  return finish_launch();
}

// thin packings for the tensor-core path of the thin layers (bf16 only):
//   pack_weight_thin : Wt[n][(kh*4+kw)*8 + c] = W[..](n, c, tap')   with tap' = flip ? 15 - tap : tap, zero for c >= C
//                      (n, c) = (d0, d1) if n_is_d0 else (d1, d0)
//   pack_weight_pad16: Wp[t][r][k], r < 16 (zero rows for r >= Nn), (n, k) = (d0, d1) if n_is_d0 else (d1, d0)
__global__ void __launch_bounds__(256)
pack_weight_thin_kernel(const float* __restrict__ w, int D0, int D1, int n_is_d0, int flip, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int Nn = n_is_d0 ? D0 : D1, C = n_is_d0 ? D1 : D0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nn * 128) return;
  const int n = i / 128, col = i % 128, tap = col / 8, c = col % 8;
  float v = 0.f;
  if (c < C) {
    const int st = flip ? 15 - tap : tap;
    const int d0 = n_is_d0 ? n : c, d1 = n_is_d0 ? c : n;
    v = w[((long long)d0 * D1 + d1) * 16 + st];
  }
  out[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
pack_weight_pad16_kernel(const float* __restrict__ w, int D0, int D1, int n_is_d0, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int Nn = n_is_d0 ? D0 : D1, K = n_is_d0 ? D1 : D0;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16LL * 16 * K) return;
  const int k = (int)(i % K), r = (int)((i / K) % 16), t = (int)(i / (16LL * K));
  float v = 0.f;
  if (r < Nn) {
    const int d0 = n_is_d0 ? r : k, d1 = n_is_d0 ? k : r;
    v = w[((long long)d0 * D1 + d1) * 16 + t];
  }
  out[i] = __float2bfloat16_rn(v);
}

int pack_weight_thin(const float* w, int D0, int D1, int n_is_d0, int flip, void* out, cudaStream_t st) {
  const int Nn = n_is_d0 ? D0 : D1, C = n_is_d0 ? D1 : D0;
  if (C > 8 || Nn < 1) return STCGAN_EINVAL;
  launch_k(pack_weight_thin_kernel, (Nn * 128 + 255) / 256, 256, 0, st, w, D0, D1, n_is_d0, flip, static_cast<__nv_bfloat16*>(out));
  return finish_launch();
}

int pack_weight_pad16(const float* w, int D0, int D1, int n_is_d0, void* out, cudaStream_t st) {
  const int Nn = n_is_d0 ? D0 : D1, K = n_is_d0 ? D1 : D0;
  if (Nn > 16 || Nn < 1) return STCGAN_EINVAL;
  const long long total = 16LL * 16 * K;
  launch_k(pack_weight_pad16_kernel, (unsigned)((total + 255) / 256), 256, 0, st, w, D0, D1, n_is_d0, static_cast<__nv_bfloat16*>(out));
  return finish_launch();
}

// G[t][d0][d1] -> grad[d0][d1][16]
__global__ void __launch_bounds__(256)
unpack_grad_kernel(const float* __restrict__ g, int D0, int D1, float* __restrict__ grad, int accumulate) {
  pdl_prologue();
  __shared__ float tile[16][32 * 8 + 1];   // [t][d0_local*32 + d1_local], 8 d0 x 32 d1 per block
  const int a0 = blockIdx.y * 8, b0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 16 * 256; i += 256) {
    const int t = i / 256, r = (i / 32) % 8, c = i % 32;
    const int d0 = a0 + r, d1 = b0 + c;
    tile[t][r * 32 + c] = (d0 < D0 && d1 < D1) ? g[((long long)t * D0 + d0) * D1 + d1] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 512; i += 256) {
    const int r = i / 512, c = (i % 512) / 16, t = i % 16;
    const int d0 = a0 + r, d1 = b0 + c;
    if (d0 < D0 && d1 < D1) {
      float* dst = &grad[((long long)d0 * D1 + d1) * 16 + t];
      const float v = tile[t][r * 32 + c];
      *dst = accumulate ? *dst + v : v;
    }
  }
}

int unpack_grad(const float* g, int D0, int D1, float* grad, int accumulate, cudaStream_t st) {
  if (D0 <= 0 || D1 <= 0) return STCGAN_EINVAL;
  dim3 grid((D1 + 31) / 32, (D0 + 7) / 8);
  launch_k(unpack_grad_kernel, grid, 256, 0, st, g, D0, D1, grad, accumulate);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------------
// module-boundary layouts
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ s1, int c1,
                  const float* __restrict__ s2, int c2, int N, int H, int W, int border, T* __restrict__ out, int Cpad) {
  pdl_prologue();
  const int HP = H + 2 * border, WP = W + 2 * border;
  const long long HW = (long long)H * W, P = (long long)N * HP * WP;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(p % WP); const long long t = p / WP;
    const int yp = (int)(t % HP); const long long n = t / HP;
    const int yy = yp - border, xx = xp - border;
    T* o = out + p * Cpad;
    int c = 0;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const long long r = (long long)yy * W + xx;
      for (int i = 0; i < c0; ++i) o[c++] = from_f32<T>(s0[(n * c0 + i) * HW + r]);
      for (int i = 0; i < c1; ++i) o[c++] = from_f32<T>(s1[(n * c1 + i) * HW + r]);
      for (int i = 0; i < c2; ++i) o[c++] = from_f32<T>(s2[(n * c2 + i) * HW + r]);
    }
    for (; c < Cpad; ++c) o[c] = from_f32<T>(0.f);
  }
}

// bf16, 8 channels per pixel (the zero-bordered thin-layer input): one 16-byte store per pixel, all source loads of a
// pixel issued before any conversion
__global__ void __launch_bounds__(256)
pack_input8_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ s1, int c1,
                   const float* __restrict__ s2, int c2, int N, int H, int W, int border, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int HP = H + 2 * border, WP = W + 2 * border;
  const long long HW = (long long)H * W, P = (long long)N * HP * WP;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(p % WP); const long long t = p / WP;
    const int yp = (int)(t % HP); const long long n = t / HP;
    const int yy = yp - border, xx = xp - border;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const long long r = (long long)yy * W + xx;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float* src = c < c0 ? s0 + (n * c0 + c) * HW : c < c0 + c1 ? s1 + (n * c1 + (c - c0)) * HW
                         : c < c0 + c1 + c2 ? s2 + (n * c2 + (c - c0 - c1)) * HW : nullptr;
        if (src) v[c] = __ldg(src + r);
      }
    }
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out + p * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int pack_input(int dtype, const float* s0, int c0, const float* s1, int c1, const float* s2, int c2,
               int N, int H, int W, int border, void* out, int Cpad, cudaStream_t st) {
  if (c0 + c1 + c2 > Cpad || c0 < 0 || c1 < 0 || c2 < 0 || border < 0) return STCGAN_EINVAL;
  const long long P = (long long)N * (H + 2 * border) * (W + 2 * border);
  if (P == 0) return 0;
  long long b = (P + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  if (dtype == STCGAN_BF16 && Cpad == 8 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    launch_k(pack_input8_kernel, (unsigned)b, 256, 0, st, s0, c0, s1, c1, s2, c2, N, H, W, border, static_cast<__nv_bfloat16*>(out));
    return finish_launch();
  }
  if (dtype == STCGAN_F32)
    launch_k(pack_input_kernel<float>, (unsigned)b, 256, 0, st, s0, c0, s1, c1, s2, c2, N, H, W, border, static_cast<float*>(out), Cpad);
  else
    launch_k(pack_input_kernel<__nv_bfloat16>, (unsigned)b, 256, 0, st, s0, c0, s1, c1, s2, c2, N, H, W, border,
                                                                  static_cast<__nv_bfloat16*>(out), Cpad);
  return finish_launch();
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_input_grad_kernel(const T* __restrict__ g, int N, long long HW, int ldg, int coff, int cn,
                         float* __restrict__ grad, int accumulate) {
  pdl_prologue();
  const long long P = (long long)N * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, r = p % HW;
    for (int c = 0; c < cn; ++c) {
      float* dst = &grad[(n * cn + c) * HW + r];
      const float v = to_f32<T>(g[p * ldg + coff + c]);
      *dst = accumulate ? *dst + v : v;
    }
  }
}

int unpack_input_grad(int dtype, const void* g, int N, int H, int W, int ldg, int coff, int cn, float* grad,
                      int accumulate, cudaStream_t st) {
  const long long HW = (long long)H * W, P = N * HW;
  if (P == 0 || cn == 0) return 0;
  long long b = (P + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  if (dtype == STCGAN_F32)
    launch_k(unpack_input_grad_kernel<float>, (unsigned)b, 256, 0, st, static_cast<const float*>(g), N, HW, ldg, coff, cn, grad, accumulate);
  else
    launch_k(unpack_input_grad_kernel<__nv_bfloat16>, (unsigned)b, 256, 0, st, static_cast<const __nv_bfloat16*>(g), N, HW, ldg,
                                                                         coff, cn, grad, accumulate);
  return finish_launch();
}

// generic NHWC <-> NCHW through a 32x32 smem transpose over (pixel, channel)
template <typename T, bool TO_NCHW>
__global__ void __launch_bounds__(256)
transpose_kernel(const void* __restrict__ src, void* __restrict__ dst, long long HW, int C, int ld) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32; const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
  if (TO_NCHW) {
    const T* x = static_cast<const T*>(src); float* o = static_cast<float*>(dst);
    for (int j = ty; j < 32; j += 8) {   // rows = pixels, tx = channel (contiguous in NHWC)
      const long long p = p0 + j; const int c = c0 + tx;
      tile[j][tx] = (p < HW && c < C) ? to_f32<T>(x[(n * HW + p) * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {   // rows = channels, tx = pixel (contiguous in NCHW)
      const int c = c0 + j; const long long p = p0 + tx;
      if (c < C && p < HW) o[(n * C + c) * HW + p] = tile[tx][j];
    }
  } else {
    const float* x = static_cast<const float*>(src); T* o = static_cast<T*>(dst);
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j; const long long p = p0 + tx;
      tile[j][tx] = (c < C && p < HW) ? x[(n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const long long p = p0 + j; const int c = c0 + tx;
      if (p < HW && c < C) o[(n * HW + p) * ld + c] = from_f32<T>(tile[tx][j]);
    }
  }
}

int nhwc_to_nchw(int dtype, const void* x, int N, int H, int W, int C, int ld, float* out, cudaStream_t st) {
  const long long HW = (long long)H * W;
  if (N == 0 || HW == 0 || C == 0) return 0;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N);
  if (dtype == STCGAN_F32) launch_k(transpose_kernel<float, true>, grid, 256, 0, st, x, out, HW, C, ld);
  else launch_k(transpose_kernel<__nv_bfloat16, true>, grid, 256, 0, st, x, out, HW, C, ld);
  return finish_launch();
}

int nchw_to_nhwc(int dtype, const float* x, int N, int H, int W, int C, void* out, int ld, cudaStream_t st) {
  const long long HW = (long long)H * W;
  if (N == 0 || HW == 0 || C == 0) return 0;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N);
  if (dtype == STCGAN_F32) launch_k(transpose_kernel<float, false>, grid, 256, 0, st, x, out, HW, C, ld);
  else launch_k(transpose_kernel<__nv_bfloat16, false>, grid, 256, 0, st, x, out, HW, C, ld);
  return finish_launch();
}

template <typename T>
__global__ void __launch_bounds__(256)
out_act_bwd_kernel(int act, const float* __restrict__ o, const float* __restrict__ d, int N, int H, int W, int C,
                   int border, T* __restrict__ g, int ldg) {
  pdl_prologue();
  const int HP = H + 2 * border, WP = W + 2 * border;
  const long long HW = (long long)H * W, P = (long long)N * HP * WP;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(p % WP); const long long t = p / WP;
    const int yp = (int)(t % HP); const long long n = t / HP;
    const int yy = yp - border, xx = xp - border;
    const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const long long r = (long long)yy * W + xx;
    for (int c = 0; c < ldg; ++c) {
      float gv = 0.f;
      if (in && c < C) {
        const float ov = o[(n * C + c) * HW + r], dv = d[(n * C + c) * HW + r];
        gv = act == STCGAN_ACT_TANH ? dv * (1.f - ov * ov) : act == STCGAN_ACT_SIGMOID ? dv * ov * (1.f - ov) : dv;
      } else if (border == 0 && c >= C) {
        continue;     // un-bordered mode keeps the historical contract: only the C real channels are written
      }
      g[p * ldg + c] = from_f32<T>(gv);
    }
  }
}

// bf16, bordered 8-channel gradient (thin tensor-core layers): one 16-byte store per pixel
__global__ void __launch_bounds__(256)
out_act_bwd8_kernel(int act, const float* __restrict__ o, const float* __restrict__ d, int N, int H, int W, int C,
                    int border, __nv_bfloat16* __restrict__ g) {
  pdl_prologue();
  const int HP = H + 2 * border, WP = W + 2 * border;
  const long long HW = (long long)H * W, P = (long long)N * HP * WP;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(p % WP); const long long t = p / WP;
    const int yp = (int)(t % HP); const long long n = t / HP;
    const int yy = yp - border, xx = xp - border;
    float gv[8], ov[8], dv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { gv[c] = 0.f; ov[c] = 0.f; dv[c] = 0.f; }
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const long long r = (long long)yy * W + xx;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) { ov[c] = __ldg(o + (n * C + c) * HW + r); dv[c] = __ldg(d + (n * C + c) * HW + r); }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        gv[c] = act == STCGAN_ACT_TANH ? dv[c] * (1.f - ov[c] * ov[c]) : act == STCGAN_ACT_SIGMOID ? dv[c] * ov[c] * (1.f - ov[c]) : dv[c];
    }
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(gv[2 * e], gv[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(g + p * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int out_act_bwd(int dtype, int act, const float* o, const float* d, int N, int H, int W, int C, int border, void* g, int ldg,
                cudaStream_t st) {
  const long long P = (long long)N * (H + 2 * border) * (W + 2 * border);
  if (P == 0) return 0;
  long long b = (P + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  if (dtype == STCGAN_BF16 && border > 0 && ldg == 8 && C <= 8 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    launch_k(out_act_bwd8_kernel, (unsigned)b, 256, 0, st, act, o, d, N, H, W, C, border, static_cast<__nv_bfloat16*>(g));
    return finish_launch();
  }
  if (dtype == STCGAN_F32) launch_k(out_act_bwd_kernel<float>, (unsigned)b, 256, 0, st, act, o, d, N, H, W, C, border, static_cast<float*>(g), ldg);
  else launch_k(out_act_bwd_kernel<__nv_bfloat16>, (unsigned)b, 256, 0, st, act, o, d, N, H, W, C, border, static_cast<__nv_bfloat16*>(g), ldg);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------------
// fused losses: value + gradient for up to 8 terms in one launch
// ---------------------------------------------------------------------------------------------
struct LossTerms { stcgan_loss_term t[8]; int n; long long first_block[9]; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int LOSS_ELEMS_PER_BLOCK = 256 * 8;

__global__ void __launch_bounds__(256)
fused_loss_kernel(const LossTerms lt, float* __restrict__ loss_out) {
  pdl_prologue();
  __shared__ float red[8];
  int ti = 0;
  while (ti + 1 < lt.n && (long long)blockIdx.x >= lt.first_block[ti + 1]) ++ti;
  const stcgan_loss_term t = lt.t[ti];
  const long long base = ((long long)blockIdx.x - lt.first_block[ti]) * LOSS_ELEMS_PER_BLOCK;
  const float inv_n = 1.f / (float)t.n;
  const float gw = t.weight * inv_n;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i >= t.n) continue;
    const float a = t.a[i];
    float term, grad;
    if (t.kind == 0) {                 // L1 (F.l1_loss: sign(a-b), 0 at equality)
      const float d = a - t.b[i];
      term = fabsf(d);
      grad = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else if (t.kind == 1) {          // MSE against a scalar label
      const float d = a - t.target;
      term = d * d;
      grad = 2.f * d;
    } else {                           // BCE with logits: max(a,0) - a*y + log1p(exp(-|a|))
      term = fmaxf(a, 0.f) - a * t.target + log1pf(expf(-fabsf(a)));
      grad = 1.f / (1.f + expf(-a)) - t.target;
    }
    s += term;
    if (t.grad) {
      const float gv = gw * grad;
      t.grad[i] = t.accumulate ? t.grad[i] + gv : gv;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(&loss_out[t.slot], v * t.loss_weight * inv_n);
  }
}

int fused_loss(const stcgan_loss_term* terms, int nterms, float* loss_out, cudaStream_t st) {
  if (nterms < 1 || nterms > 8) return STCGAN_EINVAL;
  LossTerms lt; lt.n = nterms;
  long long nb = 0;
  for (int i = 0; i < nterms; ++i) {
    if (terms[i].n <= 0 || terms[i].kind < 0 || terms[i].kind > 2 || !terms[i].a) return STCGAN_EINVAL;
    if (terms[i].kind == 0 && !terms[i].b) return STCGAN_EINVAL;
    lt.t[i] = terms[i];
    lt.first_block[i] = nb;
    nb += (terms[i].n + LOSS_ELEMS_PER_BLOCK - 1) / LOSS_ELEMS_PER_BLOCK;
  }
  lt.first_block[nterms] = nb;
  launch_k(fused_loss_kernel, (unsigned)nb, 256, 0, st, lt, loss_out);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------------
// relativistic logits (src/loss.py:88-96, 102-110): out[n,i] = a[n,i] - b[n,i]               (RpGAN)
//                                                   out[n,i] = a[n,i] - mean_n' b[n',i]      (RaGAN: batch mean, dim 0)
// and the gradient w.r.t. b of the same map: db[n,i] = -g[n,i]  /  -(1/N) sum_n' g[n',i]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rel_logits_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, long long M, int avg, int backward,
                  float* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    float mb = 0.f;
    if (avg) {
      for (int n = 0; n < N; ++n) mb += b[(long long)n * M + i];
      mb = mb / (float)N;
    }
    for (int n = 0; n < N; ++n) {
      const long long k = (long long)n * M + i;
      out[k] = backward ? -(avg ? mb : b[k]) : a[k] - (avg ? mb : b[k]);
    }
  }
}

int rel_logits(const float* a, const float* b, int N, long long M, int avg, int backward, float* out, cudaStream_t st) {
  if (N <= 0 || M <= 0) return 0;
  long long blocks = (M + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(rel_logits_kernel, (unsigned)blocks, 256, 0, st, a, b, N, M, avg, backward, out);
  return finish_launch();
}

// ---------------------------------------------------------------------------------------------
// multi-tensor Adam (torch.optim.Adam semantics: eps added after sqrt(v_hat); no weight decay / amsgrad)
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// ---------------------------------------------------------------------------------------------
constexpr int ADAM_CHUNK = 256 * 16;   // elements per block (plain / untiled tensors)
constexpr int ADAM_TILE = 32;          // tiled tensors: one block = 32 x 32 (d0, d1) pairs x 16 taps
constexpr int ADAM_G_PITCH = 17;       // floats per pair in the staged gradient tile (odd: conflict-free scalar access)
constexpr int ADAM_W_PITCH = 40;       // bf16 per row of the staged packed-weight tile (16-byte aligned rows)
constexpr int ADAM_SMEM = ADAM_TILE * ADAM_TILE * ADAM_G_PITCH * 4 + 16 * ADAM_TILE * ADAM_W_PITCH * 2;

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float beta1, float beta2, float eps,
                                            float step_size, float inv_bc2_sqrt) {
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {   // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256)
adam_kernel(const stcgan_adam_tensor* __restrict__ table, const int32_t* __restrict__ blocks, int nblocks,
            const float* __restrict__ hyper) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t adam_smem[];
  const float step_size = hyper[6], inv_bc2_sqrt = hyper[7], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3],
              grad_scale = hyper[4];
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
  __syncthreads();       // the staging tiles of the previous block of work are free
  const int ti = blocks[2 * blk], chunk = blocks[2 * blk + 1];
  const stcgan_adam_tensor t = table[ti];
  const long long base = (long long)chunk * ADAM_CHUNK;
  if (t.d0 > 0 && t.p1 != nullptr) {
    // ---- tiled: packed gradients G[tap][d0][d1], parameters [d0][d1][16], bf16 copies P1[tap][d0][d1], P2[tap][d1][d0].
    // Every global access of the block is a run of >= 64 contiguous bytes: the gradient tile and the two bf16 tiles go
    // through shared memory, p / m / v stream as float4 with consecutive threads on consecutive addresses.
    float* g_s = reinterpret_cast<float*>(adam_smem);                                    // [32*32 pairs][17]
    __nv_bfloat16* w_s = reinterpret_cast<__nv_bfloat16*>(adam_smem + ADAM_TILE * ADAM_TILE * ADAM_G_PITCH * 4);   // [16][32][40]
    const long long plane = (long long)t.d0 * t.d1;
    const int tiles1 = t.d1 / ADAM_TILE;
    const int td0 = (chunk / tiles1) * ADAM_TILE, td1 = (chunk % tiles1) * ADAM_TILE;
    // phase 1: gradient tile, 128-byte rows per (tap, d0)
    {
      float4 gr[16];          // all 16 loads of this thread in flight before the first shared-memory store
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int i = u * 256 + threadIdx.x;
        const int tap = i / (ADAM_TILE * 8), r0 = (i / 8) % ADAM_TILE, c4 = (i % 8) * 4;
        gr[u] = ldg_stream4(t.g + (long long)tap * plane + (long long)(td0 + r0) * t.d1 + td1 + c4);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int i = u * 256 + threadIdx.x;
        const int tap = i / (ADAM_TILE * 8), r0 = (i / 8) % ADAM_TILE, c4 = (i % 8) * 4;
        float* d = g_s + (r0 * ADAM_TILE + c4) * ADAM_G_PITCH + tap;
        d[0] = gr[u].x; d[ADAM_G_PITCH] = gr[u].y; d[2 * ADAM_G_PITCH] = gr[u].z; d[3 * ADAM_G_PITCH] = gr[u].w;
      }
    }
    __syncthreads();
    // phase 2: 32 rows of 32 pairs x 16 parameters = 128 float4 per row; 2 rows per pass, 4 passes in flight
#pragma unroll 1
    for (int it = 0; it < 16; it += 4) {
      float4 p[4], m[4], v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = (it + u) * 256 + threadIdx.x, r0 = idx / 128, f = idx % 128;
        const long long off = ((long long)(td0 + r0) * t.d1 + td1) * 16 + f * 4;
        p[u] = ldg_stream4(t.p + off); m[u] = ldg_stream4(t.m + off); v[u] = ldg_stream4(t.v + off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = (it + u) * 256 + threadIdx.x, r0 = idx / 128, f = idx % 128, pair = f / 4, q = f % 4;
        const long long off = ((long long)(td0 + r0) * t.d1 + td1) * 16 + f * 4;
        const float* gs = g_s + (r0 * ADAM_TILE + pair) * ADAM_G_PITCH + 4 * q;
        adam_update(p[u].x, m[u].x, v[u].x, gs[0] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
        adam_update(p[u].y, m[u].y, v[u].y, gs[1] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
        adam_update(p[u].z, m[u].z, v[u].z, gs[2] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
        adam_update(p[u].w, m[u].w, v[u].w, gs[3] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
        *reinterpret_cast<float4*>(t.p + off) = p[u];
        *reinterpret_cast<float4*>(t.m + off) = m[u];
        *reinterpret_cast<float4*>(t.v + off) = v[u];
        __nv_bfloat16* ws = w_s + ((4 * q) * ADAM_TILE + r0) * ADAM_W_PITCH + pair;
        ws[0] = __float2bfloat16_rn(p[u].x);
        ws[ADAM_TILE * ADAM_W_PITCH] = __float2bfloat16_rn(p[u].y);
        ws[2 * ADAM_TILE * ADAM_W_PITCH] = __float2bfloat16_rn(p[u].z);
        ws[3 * ADAM_TILE * ADAM_W_PITCH] = __float2bfloat16_rn(p[u].w);
      }
    }
    __syncthreads();
    // phase 3: bf16 tap-major copies, 64-byte runs: P1[tap][d0][d1 0..31], P2[tap][d1][d0 0..31]
    __nv_bfloat16* p1 = static_cast<__nv_bfloat16*>(t.p1);
    __nv_bfloat16* p2 = static_cast<__nv_bfloat16*>(t.p2);
#pragma unroll 2
    for (int i = threadIdx.x; i < 16 * ADAM_TILE * 4; i += 256) {
      const int tap = i / (ADAM_TILE * 4), r0 = (i / 4) % ADAM_TILE, ch = i % 4;
      const uint4 w = *reinterpret_cast<const uint4*>(w_s + (tap * ADAM_TILE + r0) * ADAM_W_PITCH + ch * 8);
      *reinterpret_cast<uint4*>(p1 + (long long)tap * plane + (long long)(td0 + r0) * t.d1 + td1 + ch * 8) = w;
    }
#pragma unroll 2
    for (int i = threadIdx.x; i < 16 * ADAM_TILE * 4; i += 256) {
      const int tap = i / (ADAM_TILE * 4), c0 = (i / 4) % ADAM_TILE, ch = i % 4;
      const uint16_t* src = reinterpret_cast<const uint16_t*>(w_s) + (tap * ADAM_TILE + ch * 8) * ADAM_W_PITCH + c0;
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        w[e] = (uint32_t)src[(2 * e) * ADAM_W_PITCH] | ((uint32_t)src[(2 * e + 1) * ADAM_W_PITCH] << 16);
      *reinterpret_cast<uint4*>(p2 + ((long long)tap * t.d1 + td1 + c0) * t.d0 + td0 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else if (t.d0 > 0) {
    // packed gradients without bf16 refresh (thin layers): one thread owns one (d0,d1) pair = 16 contiguous parameters
    const long long plane = (long long)t.d0 * t.d1;
    const long long r = base / 16 + threadIdx.x;
    if (r >= plane) continue;
    float* pp = t.p + r * 16; float* pm = t.m + r * 16; float* pv = t.v + r * 16;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      float p = pp[k], m = pm[k], v = pv[k];
      adam_update(p, m, v, t.g[(long long)k * plane + r] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
      pp[k] = p; pm[k] = m; pv[k] = v;
    }
  } else {
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const long long i = base + j * 256 + threadIdx.x;
      if (i >= t.n) continue;
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_update(p, m, v, t.g[i] * grad_scale, beta1, beta2, eps, step_size, inv_bc2_sqrt);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
    }
  }
  }
}

// steps_done += 1; bias corrections in double like torch (1 - beta^t)
__global__ void adam_tick_kernel(float* __restrict__ hyper) {
  pdl_prologue();
  const float step = hyper[5] + 1.f;
  hyper[5] = step;
  const double bc1 = 1.0 - pow((double)hyper[1], (double)step);
  const double bc2 = 1.0 - pow((double)hyper[2], (double)step);
  hyper[6] = (float)((double)hyper[0] / bc1);
  hyper[7] = (float)(1.0 / sqrt(bc2));
}

static int adam_update(const stcgan_adam_tensor* table, const int32_t* blocks, int nblocks, float* hyper, int tick, int max_ctas,
                       cudaStream_t st) {
  if (nblocks <= 0) return STCGAN_EINVAL;
  if (tick) {
    launch_k(adam_tick_kernel, 1, 1, 0, st, hyper);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ADAM_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int grid = (max_ctas > 0 && max_ctas < nblocks) ? max_ctas : nblocks;
  launch_k(adam_kernel, grid, 256, ADAM_SMEM, st, table, blocks, nblocks, hyper);
  return finish_launch();
}

int adam_step(const stcgan_adam_tensor* table, const int32_t* blocks, int nblocks, float* hyper, cudaStream_t st) {
  return adam_update(table, blocks, nblocks, hyper, 1, 0, st);
}

// a sub-range of the block list (the tensors of one network): `tick` advances the step counter / bias corrections and
// must be set on exactly one of the partial launches of a step -- the first one, the others must be ordered after it
int adam_step_range(const stcgan_adam_tensor* table, const int32_t* blocks, int first_block, int nblocks, float* hyper, int tick,
                    int max_ctas, cudaStream_t st) {
  if (first_block < 0) return STCGAN_EINVAL;
  return adam_update(table, blocks + 2 * (long long)first_block, nblocks, hyper, tick, max_ctas, st);
}

// ---------------------------------------------------------------------------------------------
// float2uint: numpy semantics  (np.clip(a*0.5+0.5, 0, 1) * 255).astype(uint8), all in fp32
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t f2u(float a) {
  a = fminf(fmaxf(a, 0.f), 1.f);
  return (uint8_t)(__fmul_rn(a, 255.f));   // C-style truncation toward zero, like astype(uint8) on [0,255]
}

__global__ void __launch_bounds__(256)
float2uint_kernel(const float* __restrict__ in, long long n, uint8_t* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = f2u(in[i]);
}

__global__ void __launch_bounds__(256)
float2uint_hwc_kernel(const float* __restrict__ in, int N, int C, long long HW, uint8_t* __restrict__ out) {
  pdl_prologue();
  const long long P = (long long)N * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, r = p % HW;
    for (int c = 0; c < C; ++c) {
      const float v = in[(n * C + c) * HW + r];
      out[p * C + c] = f2u(__fadd_rn(__fmul_rn(v, 0.5f), 0.5f));   // no FMA contraction: numpy rounds twice
    }
  }
}

// dataset transform on the GPU: uint8 HWC image -> float32 CHW in [-1, 1], exactly as src/utils.py:60-62 (astype(float32)/255)
// followed by src/dataset.py:152 ((s.transpose(2,0,1) - 0.5) * 2), every step rounded to float32 like numpy
__global__ void __launch_bounds__(256)
u8_to_nchw_kernel(const uint8_t* __restrict__ in, int N, int C, long long HW, float* __restrict__ out) {
  pdl_prologue();
  const long long P = (long long)N * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, r = p % HW;
    for (int c = 0; c < C; ++c) {
      const float f = __fdiv_rn((float)in[p * C + c], 255.f);
      out[(n * C + c) * HW + r] = __fmul_rn(__fsub_rn(f, 0.5f), 2.f);
    }
  }
}

int u8_to_nchw(const uint8_t* in, int N, int H, int W, int C, float* out, cudaStream_t st) {
  const long long HW = (long long)H * W, P = N * HW;
  if (P == 0 || C == 0) return 0;
  long long b = (P + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  launch_k(u8_to_nchw_kernel, (unsigned)b, 256, 0, st, in, N, C, HW, out);
  return finish_launch();
}

int float2uint(const float* in, long long n, uint8_t* out, cudaStream_t st) {
  if (n == 0) return 0;
  long long b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  launch_k(float2uint_kernel, (unsigned)b, 256, 0, st, in, n, out);
  return finish_launch();
}

int float2uint_hwc(const float* in, int N, int C, int H, int W, uint8_t* out, cudaStream_t st) {
  const long long HW = (long long)H * W, P = N * HW;
  if (P == 0) return 0;
  long long b = (P + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  launch_k(float2uint_hwc_kernel, (unsigned)b, 256, 0, st, in, N, C, HW, out);
  return finish_launch();
}

}  // namespace stcgan

// CUDA-core (FFMA, fp32 accumulate) tap-GEMM convolution kernels.
//
// These serve (i) the fp32 parity mode of every layer, and (ii) in bf16 mode the thin layers whose
// GEMMs cannot fill a tcgen05 tile (Cin in {3,4,7} first layers, Cout in {1,3} last layers) and are
// HBM-bound anyway (SURVEY.md section 7 "Hard parts").  Any shape, any pitch, both element types.
//
// forward / dgrad:  out[p, n] = sum_taps sum_k in[gather(p), k] * Wp[t][n][k]      (tapconv_ffma)
// wgrad:            G[t][d0][d1] += sum_p S[p, d0] * L[win_t(p), d1]               (tapwgrad_ffma)
#include "common.cuh"

namespace stcgan {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT)
tapconv_ffma_kernel(const Geom g, const T* __restrict__ x, int K, int ldx,
                    const T* __restrict__ wp, const float* __restrict__ bias, int act,
                    void* __restrict__ yv, int Nout, int ldy, int out_nchw_f32) {
  pdl_prologue();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int cls = blockIdx.z;
  const int GH = g.grid_h(cls), GW = g.grid_w(cls);
  const long long M = (long long)g.N * GH * GW;
  const long long m0 = (long long)blockIdx.x * BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * BN;

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // 16 x 16 threads, 4x4 outputs each

  // loader role: row = tid / 4 (pixel for A, out-channel for B), k-quad = tid % 4
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long pm = m0 + lrow;
  int pn = 0, pa = 0, pb = 0;
  const bool prow_ok = pm < M;
  if (prow_ok) {
    pn = (int)(pm / ((long long)GH * GW));
    const int r = (int)(pm % ((long long)GH * GW));
    pa = r / GW; pb = r % GW;
  }
  const int wn = n0 + lrow;
  const bool wrow_ok = wn < Nout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int j = 0; j < g.ntaps; ++j) {
    const Tap tp = g.tap[cls][j];
    const int iy = pa * g.istride + tp.dy, ix = pb * g.istride + tp.dx;
    const bool in_ok = prow_ok && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
    const T* xrow = x + ((long long)(pn * g.IH + (in_ok ? iy : 0)) * g.IW + (in_ok ? ix : 0)) * ldx;
    const T* wrow = wp + ((long long)tp.wtap * Nout + (wrow_ok ? wn : 0)) * K;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + lk + q;
        As[lk + q][lrow] = (in_ok && k < K) ? to_f32<T>(xrow[k]) : 0.f;
        Bs[lk + q][lrow] = (wrow_ok && k < K) ? to_f32<T>(wrow[k]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], bv[jj], acc[i][jj]);
      }
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int n = (int)(m / ((long long)GH * GW));
    const int r = (int)(m % ((long long)GH * GW));
    const int oy = (r / GW) * g.ostride + g.oy0[cls], ox = (r % GW) * g.ostride + g.ox0[cls];
    if (oy >= g.OH || ox >= g.OW) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int co = n0 + tx * 4 + jj;
      if (co >= Nout) continue;
      float v = acc[i][jj] + (bias ? bias[co] : 0.f);
      v = act_fwd(act, v);
      if (out_nchw_f32) {
        static_cast<float*>(yv)[(((long long)n * Nout + co) * g.OH + oy) * g.OW + ox] = v;
      } else {
        static_cast<T*>(yv)[((long long)(n * g.OH + oy) * g.OW + ox) * ldy + co] = from_f32<T>(v);
      }
    }
  }
}

template <typename T>
static int launch_tapconv_ffma(const Geom& g, const void* x, int K, int ldx, const void* wp, const float* bias,
                               int act, void* y, int Nout, int ldy, int out_nchw_f32, cudaStream_t st) {
  long long maxM = 0;
  for (int c = 0; c < g.nclass; ++c) {
    long long m = (long long)g.N * g.grid_h(c) * g.grid_w(c);
    if (m > maxM) maxM = m;
  }
  if (maxM == 0 || Nout == 0) return 0;
  dim3 grid((unsigned)((maxM + BM - 1) / BM), (unsigned)((Nout + BN - 1) / BN), (unsigned)g.nclass);
  launch_k(tapconv_ffma_kernel<T>, grid, NT, 0, st, g, static_cast<const T*>(x), K, ldx, static_cast<const T*>(wp), bias,
                                              act, y, Nout, ldy, out_nchw_f32);
  return finish_launch();
}

int tapconv_ffma(const Geom& g, int dtype, const void* x, int K, int ldx, const void* wp, const float* bias, int act,
                 void* y, int Nout, int ldy, int out_nchw_f32, cudaStream_t st) {
  if (dtype == STCGAN_F32) return launch_tapconv_ffma<float>(g, x, K, ldx, wp, bias, act, y, Nout, ldy, out_nchw_f32, st);
  return launch_tapconv_ffma<__nv_bfloat16>(g, x, K, ldx, wp, bias, act, y, Nout, ldy, out_nchw_f32, st);
}

// ------------------------------------------------------------------------------------------------
// wgrad: one block = (64 d0 x 64 d1) tile of one tap, over a slice of the pixel range; atomics combine slices
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
tapwgrad_ffma_kernel(int N, int SH, int SW, int D0, int lds, int LH, int LW, int D1, int ldl, int stride,
                     const T* __restrict__ S, const T* __restrict__ L, float* __restrict__ G,
                     int tiles1, long long pix_per_split) {
  pdl_prologue();
  __shared__ float Ss[BK][BM + 4];   // [pixel][d0]
  __shared__ float Ls[BK][BN + 4];   // [pixel][d1]

  const int t = blockIdx.y;           // tap = kh*4 + kw
  const int kh = t / 4, kw = t % 4;
  const int a0 = (blockIdx.x / tiles1) * BM, b0 = (blockIdx.x % tiles1) * BN;
  const long long P = (long long)N * SH * SW;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > P) p_end = P;
  if (p_begin >= p_end) return;

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int lp = tid / 16, lc = (tid % 16) * 4;   // loader: pixel lp (0..15), channel quad lc (0..60)

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pc = p_begin; pc < p_end; pc += BK) {
    const long long p = pc + lp;
    const bool pok = p < p_end;
    int n = 0, sy = 0, sx = 0;
    if (pok) {
      n = (int)(p / ((long long)SH * SW));
      const int r = (int)(p % ((long long)SH * SW));
      sy = r / SW; sx = r % SW;
    }
    const int ly = sy * stride - 1 + kh, lx = sx * stride - 1 + kw;
    const bool lok = pok && ly >= 0 && ly < LH && lx >= 0 && lx < LW;
    const T* srow = S + p * (long long)lds;
    const T* lrow = L + ((long long)(n * LH + (lok ? ly : 0)) * LW + (lok ? lx : 0)) * ldl;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c0 = a0 + lc + q, c1 = b0 + lc + q;
      Ss[lp][lc + q] = (pok && c0 < D0) ? to_f32<T>(srow[c0]) : 0.f;
      Ls[lp][lc + q] = (lok && c1 < D1) ? to_f32<T>(lrow[c1]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Ss[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ls[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], bv[jj], acc[i][jj]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c0 = a0 + ty * 4 + i;
    if (c0 >= D0) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c1 = b0 + tx * 4 + jj;
      if (c1 >= D1) continue;
      atomicAdd(&G[((long long)t * D0 + c0) * D1 + c1], acc[i][jj]);
    }
  }
}

int tapwgrad_ffma(int geom, int dtype, const void* S, int N, int SH, int SW, int D0, int lds,
                  const void* L, int LH, int LW, int D1, int ldl, float* G, cudaStream_t st) {
  const int stride = geom == STCGAN_GEOM_WIN_S2 ? 2 : 1;
  const long long P = (long long)N * SH * SW;
  if (P == 0 || D0 == 0 || D1 == 0) return 0;
  const int tiles0 = (D0 + BM - 1) / BM, tiles1 = (D1 + BN - 1) / BN;
  // split the pixel range so that the grid has at least ~4 waves of 148 SMs, chunks >= 256 pixels
  long long want = (4LL * 148 + (long long)tiles0 * tiles1 * 16 - 1) / ((long long)tiles0 * tiles1 * 16);
  long long max_split = (P + 255) / 256;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  long long pps = (P + want - 1) / want;
  pps = (pps + BK - 1) / BK * BK;
  const unsigned splits = (unsigned)((P + pps - 1) / pps);
  dim3 grid((unsigned)(tiles0 * tiles1), 16, splits);
  if (dtype == STCGAN_F32)
    launch_k(tapwgrad_ffma_kernel<float>, grid, NT, 0, st, N, SH, SW, D0, lds, LH, LW, D1, ldl, stride,
                                                     static_cast<const float*>(S), static_cast<const float*>(L), G, tiles1, pps);
  else
    launch_k(tapwgrad_ffma_kernel<__nv_bfloat16>, grid, NT, 0, st, N, SH, SW, D0, lds, LH, LW, D1, ldl, stride,
                                                             static_cast<const __nv_bfloat16*>(S),
                                                             static_cast<const __nv_bfloat16*>(L), G, tiles1, pps);
  return finish_launch();
}

}  // namespace stcgan

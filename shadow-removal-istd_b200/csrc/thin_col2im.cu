// Thin-N convolutions (at most 8 output channels) as ONE pixel GEMM + an in-CTA col2im.
//
// The layers with a thin output -- ConvTranspose2d 128 -> {1,3} (stcgan_g.py:93-95), Conv2d 512 -> 1 stride 1
// (stcgan_d.py:49-50) and the input gradient of the first Conv2d layers (Cin in {3,4,7}) -- are HBM-bound: their cost is
// reading the fat tensor.  The tap-GEMM formulation reads every input pixel once per tap (16x or 4x, from L2); here the
// taps move into the GEMM's N dimension instead:
//
//     Pm[pixel][(tap, c)] = sum_k x[pixel][k] * Wt[(tap, c)][k]          (M = 128 input pixels, N = 16 * cpad, K = Cin)
//
// so the fat tensor is read once (plus a one-pixel halo), and the 16 partial planes are summed inside the CTA:
//     mode 0 (stride-2 scatter: ConvTranspose2d forward / Conv2d-s2 input gradient)
//         out[2i-1+kh, 2j-1+kw, c] = sum over the (i, j, kh, kw) that hit it of Pm[(i,j)][(kh,kw,c)]     (4 terms)
//     mode 1 (stride-1 gather: Conv2d(k4,s1,p1) forward)
//         out[oy, ox, c] = sum_{kh,kw} Pm[(oy-1+kh, ox-1+kw)][(kh,kw,c)]                                  (16 terms)
// Tiles of th x tw input pixels overlap by one (mode 0) or three (mode 1) rows / columns so that every output pixel is
// complete inside exactly one CTA: no atomics, bias + Tanh / Sigmoid fused, NCHW fp32 or 8-channel NHWC bf16 stores.
//
// Same warp roles as tapconv_tc.cu: warp 0 TMA producer, warp 1 tcgen05.mma issuer (accumulator 128 x N fp32 in TMEM),
// warps 2-5 epilogue.
#include <cstring>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace stcgan {

struct alignas(64) PixGemmParams {
  CUtensorMap amap;        // x [NB, H, W, K], box (64, tw, th, 1); out-of-range pixels read zeros
  CUtensorMap bmap;        // Wt [16 * cpad][K], box (64, 16 * cpad)
  int kchunks, mode;
  int tw, th, tiles_w, tiles_h;
  int OH, OW, NB;
  int cpad, cout, act;
  const float* bias;
  float* y32;              // NCHW fp32 [NB, cout, OH, OW] (bias + activation), or
  __nv_bfloat16* y8;       // NHWC bf16, 8 channels per pixel (channels >= cout are written as zeros), pitch ldy
  int ldy;
  uint8_t* u8;             // optional, next to (or instead of) y32 in mode 0: NHWC uint8 [NB, OH, OW, cout] =
                           // utils.float2uint(act(v) * 0.5 + 0.5), the image CGAN.infer writes (src/cgan.py:441-446)
};

// (np.clip(a * 0.5 + 0.5, 0, 1) * 255).astype(uint8) in float32 like numpy: two roundings (no FMA), truncation
__device__ __forceinline__ uint8_t quantise_u8(float a) {
  float t = __fadd_rn(__fmul_rn(a, 0.5f), 0.5f);
  t = fminf(fmaxf(t, 0.f), 1.f);
  return (uint8_t)(__fmul_rn(t, 255.f));
}

template <int NN, int STAGES>
struct PixSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = NN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int P_PITCH = NN + 1;                        // floats per pixel row of the staged Pm tile
  static constexpr int STAGING = 128 * P_PITCH * 4;
  static constexpr int BAR_OFFSET = ((RING > STAGING ? RING : STAGING) + 15) / 16 * 16;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = NN < 32 ? 32 : NN;
};

__device__ __forceinline__ float thin_act(int act, float v) {
  if (act == STCGAN_ACT_TANH) return tanhf(v);
  if (act == STCGAN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return act_piecewise(v, act_slope(act));
}

template <int NN, int STAGES>
__global__ void __launch_bounds__(192)
pixgemm_col2im_kernel(const __grid_constant__ PixGemmParams P) {
  pdl_trigger();
  using SM = PixSmem<NN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x % P.tiles_w;
  const int ty = (blockIdx.x / P.tiles_w) % P.tiles_h;
  const int n = blockIdx.x / (P.tiles_w * P.tiles_h);
  const int halo = P.mode == 0 ? 1 : 3;
  const int r0 = ty * (P.th - halo) - 1, c0 = tx * (P.tw - halo) - 1;     // first input row / column of this tile
  const int iters = P.kchunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.amap);
    tma_prefetch_desc(&P.bmap);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
      tma_load_4d(&P.amap, &full_bar[s], a_dst, it * 64, c0, r0, n);
      tma_load_2d(&P.bmap, &full_bar[s], a_dst + SM::A_BYTES, it * 64, 0);
    }
  } else if (threadIdx.x == 32) {
    constexpr uint32_t idesc = make_idesc(128, NN, 0, 0);
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
      const uint32_t b_addr = a_addr + SM::A_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                  (it | k) != 0);
      umma_commit(&empty_bar[s]);
    }
    umma_commit(tmem_full);
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // accumulator row = input pixel (lh * tw + lw) of the tile
    float* Ps = reinterpret_cast<float*>(smem);    // [128][P_PITCH], aliases the (drained) pipeline ring
    mbar_wait_warp(tmem_full, 0, lane);
    tc_fence_after();
    if constexpr (NN == 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
      for (int c = 0; c < 16; ++c) Ps[row * SM::P_PITCH + c] = __uint_as_float(r[c]);
    } else {
#pragma unroll 1
      for (int cb = 0; cb < NN; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, r);
#pragma unroll
        for (int c = 0; c < 32; ++c) Ps[row * SM::P_PITCH + cb + c] = __uint_as_float(r[c]);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int et = threadIdx.x - 64;
    const int tw = P.tw, cpad = P.cpad;
    if (P.mode == 0) {
      const int OHt = 2 * P.th - 2, OWt = 2 * tw - 2, oy_base = 2 * r0 + 1, ox_base = 2 * c0 + 1;
      if (P.y32 || P.u8) {
        const int total = P.cout * OHt * OWt;
        for (int idx = et; idx < total; idx += 128) {
          const int lx = idx % OWt, ly = (idx / OWt) % OHt, co = idx / (OWt * OHt);
          const int oy = oy_base + ly, ox = ox_base + lx;
          if (oy < 0 || oy >= P.OH || ox < 0 || ox >= P.OW) continue;
          const int li = ly >> 1, lj = lx >> 1, khA = 2 + (ly & 1), khB = ly & 1, kwA = 2 + (lx & 1), kwB = lx & 1;
          const float* p00 = Ps + (li * tw + lj) * SM::P_PITCH + co;
          const float* p10 = p00 + tw * SM::P_PITCH;
          float v = p00[(khA * 4 + kwA) * cpad] + p00[SM::P_PITCH + (khA * 4 + kwB) * cpad] +
                    p10[(khB * 4 + kwA) * cpad] + p10[SM::P_PITCH + (khB * 4 + kwB) * cpad];
          if (P.bias) v += __ldg(P.bias + co);
          const float a = thin_act(P.act, v);
          if (P.y32) P.y32[(((long long)n * P.cout + co) * P.OH + oy) * P.OW + ox] = a;
          if (P.u8) P.u8[(((long long)n * P.OH + oy) * P.OW + ox) * P.cout + co] = quantise_u8(a);
        }
      } else {
        const int total = OHt * OWt;
        for (int idx = et; idx < total; idx += 128) {
          const int lx = idx % OWt, ly = idx / OWt;
          const int oy = oy_base + ly, ox = ox_base + lx;
          if (oy < 0 || oy >= P.OH || ox < 0 || ox >= P.OW) continue;
          const int li = ly >> 1, lj = lx >> 1, khA = 2 + (ly & 1), khB = ly & 1, kwA = 2 + (lx & 1), kwB = lx & 1;
          const float* p00 = Ps + (li * tw + lj) * SM::P_PITCH;
          const float* p10 = p00 + tw * SM::P_PITCH;
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            v[c] = c < P.cout ? p00[(khA * 4 + kwA) * cpad + c] + p00[SM::P_PITCH + (khA * 4 + kwB) * cpad + c] +
                                p10[(khB * 4 + kwA) * cpad + c] + p10[SM::P_PITCH + (khB * 4 + kwB) * cpad + c]
                              : 0.f;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(P.y8 + (((long long)n * P.OH + oy) * P.OW + ox) * P.ldy) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    } else {
      const int OHt = P.th - 3, OWt = tw - 3, oy_base = r0 + 1, ox_base = c0 + 1;
      const int total = P.cout * OHt * OWt;
      for (int idx = et; idx < total; idx += 128) {
        const int lx = idx % OWt, ly = (idx / OWt) % OHt, co = idx / (OWt * OHt);
        const int oy = oy_base + ly, ox = ox_base + lx;
        if (oy < 0 || oy >= P.OH || ox < 0 || ox >= P.OW) continue;
        float v = 0.f;
#pragma unroll
        for (int kh = 0; kh < 4; ++kh)
#pragma unroll
          for (int kw = 0; kw < 4; ++kw)
            v += Ps[((ly + kh) * tw + lx + kw) * SM::P_PITCH + (kh * 4 + kw) * cpad + co];
        if (P.bias) v += __ldg(P.bias + co);
        P.y32[(((long long)n * P.cout + co) * P.OH + oy) * P.OW + ox] = thin_act(P.act, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<SM::TMEM_COLS>(tmem_base);
  }
}

template <int NN, int STAGES>
static int launch_pixgemm(const PixGemmParams& P, unsigned grid, cudaStream_t st) {
  using SM = PixSmem<NN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pixgemm_col2im_kernel<NN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(pixgemm_col2im_kernel<NN, STAGES>, grid, 192, SM::TOTAL, st, P);
  return finish_launch();
}

// mode 0: stride-2 scatter (OH <= 2*IH + 1, OW <= 2*IW + 1), mode 1: stride-1 gather (OH <= IH - 1, OW <= IW - 1)
// wt: packed weights [(tap * cpad + c)][K] bf16 (stcgan_pack_weight_tapn), cpad in {1, 4, 8}, cout <= cpad
int thin_col2im_tc(int mode, const void* x, int NB, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                   const float* bias, int act, float* y32, void* y8, int ldy, int OH, int OW, cudaStream_t st, uint8_t* u8) {
  if (K % 64 != 0 || ldx % 8 != 0 || !al16(x) || !al16(wt)) return STCGAN_EUNSUPPORTED;
  if ((cpad != 1 && cpad != 4 && cpad != 8) || cout < 1 || cout > cpad) return STCGAN_EINVAL;
  if (u8 ? (mode != 0 || y8 != nullptr) : ((y32 == nullptr) == (y8 == nullptr))) return STCGAN_EINVAL;
  if (y8 && (mode != 0 || bias || act != STCGAN_ACT_NONE || ldy % 8 != 0 || ldy < 8 || !al16(y8))) return STCGAN_EUNSUPPORTED;
  if (mode != 0 && mode != 1) return STCGAN_EINVAL;
  PixGemmParams P;
  memset(&P, 0, sizeof(P));
  // tile shape: th x tw = 128 input pixels minimising the number of CTAs
  long long best = -1;
  for (int tw = 4; tw <= 32; tw *= 2) {
    const int th = 128 / tw;
    long long th_n, tw_n;
    if (mode == 0) { th_n = (OH + 1 + 2 * (th - 1) - 1) / (2 * (th - 1)); tw_n = (OW + 1 + 2 * (tw - 1) - 1) / (2 * (tw - 1)); }
    else           { th_n = (OH + th - 4) / (th - 3); tw_n = (OW + tw - 4) / (tw - 3); }
    const long long tiles = th_n * tw_n;
    if (best < 0 || tiles < best || (tiles == best && tw > P.tw)) { best = tiles; P.tw = tw; P.th = th; P.tiles_w = (int)tw_n; P.tiles_h = (int)th_n; }
  }
  P.kchunks = K / 64; P.mode = mode; P.OH = OH; P.OW = OW; P.NB = NB; P.cpad = cpad; P.cout = cout; P.act = act;
  P.bias = bias; P.y32 = y32; P.y8 = static_cast<__nv_bfloat16*>(y8); P.ldy = ldy; P.u8 = u8;
  int rc = encode_nhwc(&P.amap, x, K, IW, IH, NB, ldx, (long long)IW * ldx, (long long)IH * IW * ldx, P.tw, P.th, 1);
  if (rc) return rc;
  const int NN = 16 * cpad;
  rc = encode_2d(&P.bmap, wt, K, NN, NN);
  if (rc) return rc;
  const unsigned grid = (unsigned)((long long)P.tiles_w * P.tiles_h * NB);
  if (grid == 0) return 0;
  if (NN == 16) return launch_pixgemm<16, 4>(P, grid, st);
  if (NN == 64) return launch_pixgemm<64, 2>(P, grid, st);
  return launch_pixgemm<128, 2>(P, grid, st);
}

// Wt[(t * cpad + r)][k] = W(.., t) with (r, k) = (d0, d1) if n_is_d0 else (d1, d0); rows r >= Nn are zero
__global__ void __launch_bounds__(256)
pack_weight_tapn_kernel(const float* __restrict__ w, int D0, int D1, int n_is_d0, int cpad, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int Nn = n_is_d0 ? D0 : D1, K = n_is_d0 ? D1 : D0;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16LL * cpad * K) return;
  const int k = (int)(i % K), r = (int)((i / K) % cpad), t = (int)(i / ((long long)cpad * K));
  float v = 0.f;
  if (r < Nn) {
    const int d0 = n_is_d0 ? r : k, d1 = n_is_d0 ? k : r;
    v = w[((long long)d0 * D1 + d1) * 16 + t];
  }
  out[i] = __float2bfloat16_rn(v);
}

int pack_weight_tapn(const float* w, int D0, int D1, int n_is_d0, int cpad, void* out, cudaStream_t st) {
  const int Nn = n_is_d0 ? D0 : D1, K = n_is_d0 ? D1 : D0;
  if (Nn > cpad || Nn < 1 || (cpad != 1 && cpad != 4 && cpad != 8)) return STCGAN_EINVAL;
  const long long total = 16LL * cpad * K;
  launch_k(pack_weight_tapn_kernel, (unsigned)((total + 255) / 256), 256, 0, st, w, D0, D1, n_is_d0, cpad, static_cast<__nv_bfloat16*>(out));
  return finish_launch();
}

}  // namespace stcgan

// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 operands, fp32 accumulate).
//
//   forward + dgrad ("tap GEMM"):   out[p, n] = sum_taps sum_k  A_tap[p, k] * Wp[tap][n][k]
//       A_tap is a TMA box of the NHWC activation tensor, shifted by the tap offset; stride-2 windows are read
//       through four parity views of the tensor (plain tiled TMA, unit element strides); out-of-range rows,
//       columns and images are zero-filled by TMA, which implements the conv padding (and the reference's
//       odd-size F.pad) for free.  Both operands are K-major, 128-byte swizzled.
//   wgrad:                          G[tap][d0][d1] += sum_p S[p, d0] * L[win_tap(p), d1]
//       the reduction runs over pixels, which is the slow dimension of NHWC, so both operands are MN-major
//       (128-byte swizzled rows of 64 channels); split over pixel tiles, fp32 red.global.add into G.
//
// One CTA = one 128 x BN accumulator tile in TMEM.  Warp 0: TMA producer (one thread).  Warp 1: TMEM allocator and
// MMA issuer (one thread issues tcgen05.mma / tcgen05.commit).  Warps 2-5: epilogue (tcgen05.ld 32x32b, bias +
// activation, 128-bit stores).  smem ring of STAGES x (A tile + B tile) guarded by full/empty mbarriers.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace stcgan {

// ---------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ---------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 64;

struct alignas(64) TapGemmParams {
  CUtensorMap amap[4];     // activation views (view 0 only for unit-stride geometries)
  CUtensorMap bmap;        // packed weights as a 2D tensor [16*Nout rows][K]
  // per-class tap tables, already translated to view coordinates
  int8_t tdy[4][16], tdx[4][16], tview[4][16], twt[4][16];
  int8_t oy0[4], ox0[4];
  int ntaps, kchunks;
  int N, OH, OW, ostride;
  int GH, GW;              // largest class grid
  int wt, ht, nt;          // tile box: wt*ht*nt = 128 grid pixels
  int tiles_w, tiles_h;    // tiles along the grid; blockIdx.x enumerates (n-tile, h-tile, w-tile)
  int Nout, ldy, act;
  const float* bias;
  __nv_bfloat16* y;
  // thin variants
  int thin_k;              // 1: A rows are im2col windows of a zero-bordered 8-channel tensor (5D map in amap[0]),
                           //    K = 128 = two 64-wide chunks (kh pairs); ntaps = 1, kchunks = 2
  int nout_real;           // BN == 16 only: number of valid output channels (<= 16)
  float* y32;              // BN == 16 only: NCHW fp32 output (bias + activation incl. tanh / sigmoid) instead of y
  // split-K over taps (deep-K layers with few output tiles): blockIdx.z = class * ksplit + part; fp32 partial sums are
  // reduced with vector red.global.add into `part_out` [N, OH, OW, Nout_total] (zeroed by the host wrapper)
  int ksplit;
  float* part_out;
  int part_ld;
  long long* dbg;          // optional per-CTA timestamps (8 x int64 per CTA), profiling aid (STCGAN_TC_DEBUG_TIMES)
  // BatchNorm statistics fused into the epilogue: per output channel sum and sum of squares of the bf16-rounded outputs
  // (fp32 over the 128 rows of a tile, then fp64 atomics) into bn_acc[slot][2][bn_c], slot = CTA index % STCGAN_BN_SLOTS
  double* bn_acc;
  int bn_c;
  int dbg_mode;            // timing experiments only (STCGAN_TC_DBGMODE): bit 0 = skip the MMAs, bit 1 = skip the TMA loads,
                           // bit 2 = fetch the weight tile as ONE contiguous bulk copy (wrong data: timing of a tile-major layout)
  const void* wp_raw;
  // extended epilogue (inference: eval-mode BatchNorm folded into the convolution; dual activation of the U-Net skip):
  //   v = acc * ep_scale[c] + bias[c]   (ep_scale == nullptr: 1)   ->  y = act(v),  y2 = act2(v)  (y2 == nullptr: none)
  // OH / OW above are the STORE extent (a cropped destination has OH, OW smaller than the layer's true output size)
  const float* ep_scale;
  __nv_bfloat16* y2;
  int ldy2, act2;
};


__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_T(slot) do { if (P.dbg) P.dbg[(((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime(); } while (0)

// MT = number of 128-row accumulators per CTA (M tile = MT x 128 pixels): MT = 2 halves the CTA count of launches that
// would need more than one wave, shares every weight tile between two MMAs and amortises the issuing thread's per-stage
// wait + commit (~250 cycles, profiles/r01_pipeline_microbenchmarks.txt) over 8 MMAs instead of 4
// NI = number of MMA issuer threads (1, or 2 = a second issuer in a seventh warp).  The issuers alternate ring stages and
// accumulate into SEPARATE TMEM accumulators that the epilogue adds up, so no ordering between their MMAs is needed: while
// one issuer sits in its wait + commit (~250 cycles per stage, not hidden behind the 1-2 MMAs the pipe queues) the other
// one's MMAs keep the tensor pipe busy (tools/mma_pipe.cu: 524 -> 277 cycles per 128x128x64 stage from one CTA)
template <int BN, int STAGES, int MT = 1, int NI = 1>
struct TapGemmSmem {
  static constexpr int A_BYTES = MT * TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = NI * MT * BN < 32 ? 32 : NI * MT * BN;
  static_assert(NI == 1 || (BN >= 64 && NI * MT * BN <= 512 && STAGES >= 2), "two issuers need 2 x MT x BN TMEM columns");
  // epilogue reuse of the (idle) pipeline smem: bf16 staging tile, then the column-statistics scratch
  static constexpr int STAT_OFFSET = (MT * TC_BM * (BN * 2 + 16) + 127) / 128 * 128;
  static_assert(BN < 64 || STAT_OFFSET + 2 * 1024 * 4 <= BAR_OFFSET, "statistics scratch must fit in the pipeline smem");
  static_assert(MT == 1 || (BN >= 64 && MT * BN <= 512), "two accumulators need 2 x BN TMEM columns");
};

// EP = extended epilogue (per-channel scale, second output; inference only): a separate instantiation, so that the training
// kernels keep the plain epilogue (the generalised one cost 3-7 % on every launch and 33 % on the epilogue-bound thin-K
// persistent kernel when it was a run-time switch: registers and a second loop level around the TMEM reads)
template <int BN, int STAGES, int MT = 1, int NI = 1, bool EP = false>
__global__ void __launch_bounds__(NI == 2 ? 224 : 192)
tapgemm_tc_kernel(const __grid_constant__ TapGemmParams P) {
  pdl_trigger();
  using SM = TapGemmSmem<BN, STAGES, MT, NI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksplit = P.ksplit > 1 ? P.ksplit : 1;
  const int cls = blockIdx.z / ksplit, kpart = blockIdx.z % ksplit;
  const int n_col0 = blockIdx.y * BN;
  const int tw = blockIdx.x % P.tiles_w;
  const int th = (blockIdx.x / P.tiles_w) % P.tiles_h;
  const int tn = blockIdx.x / (P.tiles_w * P.tiles_h);
  const int b0 = tw * P.wt, a0 = th * P.ht, n0 = tn * P.nt;
  const int taps_here = P.ntaps / ksplit, tap0 = kpart * taps_here;
  const int iters = taps_here * P.kchunks;

  if (threadIdx.x == 0) {
    DBG_T(0);
    tma_prefetch_desc(&P.bmap);
    tma_prefetch_desc(&P.amap[0]);
    tma_prefetch_desc(&P.amap[1]);
    tma_prefetch_desc(&P.amap[2]);
    tma_prefetch_desc(&P.amap[3]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, NI);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is independent of the preceding kernel; global memory is touched only below
  if (threadIdx.x == 0) DBG_T(1);

  if (threadIdx.x == 0) {
    // ===== TMA producer =====
    // One thread feeds the whole CTA, so its per-stage instruction chain IS the pipeline's upper rate: ring slot / phase
    // are carried incrementally (no divisions) and the tap table entries of the next tap are fetched one tap ahead.
    int s = 0; uint32_t ph = 0;
    if (P.thin_k) {
      for (int kc = 0; kc < iters; ++kc) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
        if (P.dbg_mode & 2) { mbar_arrive(&full_bar[s]); }
        else {
          mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
          // one 64-byte line per (pixel, kernel row): TMA gives every inner line its own swizzle-span row, so the two
          // kernel rows of this K chunk are two consecutive [128 px][64 B] SWIZZLE_64B blocks (kernel row = slowest box dim)
          tma_load_5d(&P.amap[0], &full_bar[s], a_dst, 0, b0, a0, n0, 2 * kc);   // both kernel rows in one instruction
          tma_load_2d(&P.bmap, &full_bar[s], a_dst + SM::A_BYTES, kc * TC_BK, n_col0);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    } else {
      const int kchunks = P.kchunks;
      int view = P.tview[cls][tap0], cx = b0 + P.tdx[cls][tap0], cy = a0 + P.tdy[cls][tap0];
      int wrow = (int)P.twt[cls][tap0] * P.Nout + n_col0;
      for (int j = 0; j < taps_here; ++j) {
        const CUtensorMap* am = &P.amap[view];
        const int cxj = cx, cyj = cy, wrowj = wrow;
        if (j + 1 < taps_here) {      // next tap's table entries, off the critical path
          const int jn = tap0 + j + 1;
          view = P.tview[cls][jn]; cx = b0 + P.tdx[cls][jn]; cy = a0 + P.tdy[cls][jn];
          wrow = (int)P.twt[cls][jn] * P.Nout + n_col0;
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
          if (P.dbg_mode & 2) { mbar_arrive(&full_bar[s]); }
          else {
            mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
            tma_load_4d(am, &full_bar[s], a_dst, kc * TC_BK, cxj, cyj, n0);
            if (P.dbg_mode & 4)
              bulk_load_1d(a_dst + SM::A_BYTES, static_cast<const uint8_t*>(P.wp_raw) + (size_t)((j * kchunks + kc) % 32) * SM::B_BYTES,
                           SM::B_BYTES, &full_bar[s]);
            else
              tma_load_2d(&P.bmap, &full_bar[s], a_dst + SM::A_BYTES, kc * TC_BK, wrowj);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (threadIdx.x == 32 || (NI == 2 && threadIdx.x == 192)) {
    // ===== MMA issuer(s): issuer `who` takes the ring stages who, who + NI, ... and owns accumulator set `who` =====
    constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
    const int who = threadIdx.x == 32 ? 0 : 1;
    int s = who; uint32_t ph = 0;                 // (NI <= STAGES)
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t acc_base = tmem_base + (uint32_t)(who * MT * BN);
    const bool thin_k = P.thin_k != 0, no_mma = (P.dbg_mode & 1) != 0;
    for (int it = who; it < iters; it += NI) {
      mbar_wait(&full_bar[s], ph);
      if (it == 0) DBG_T(2);
      tc_fence_after();
      const uint32_t a_addr = smem_base + s * SM::STAGE_BYTES;
      const uint32_t b_addr = a_addr + SM::A_BYTES;
      const int first = it - who;                 // 0 on this issuer's first stage: its accumulators start from zero
      if (no_mma) { mbar_arrive(&empty_bar[s]); }
      else {
        if (thin_k) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // block k/2 (kernel row), 16 K-elements k%2 inside its 64-byte rows
            const uint64_t ad = make_smem_desc(a_addr + (k >> 1) * 8192 + (k & 1) * 32, 16, 512, 4);
            const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(acc_base, ad, bd, idesc, (first | k) != 0);
          }
        } else {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t ad = make_smem_desc(a_addr + mt * (TC_BM * TC_BK * 2) + k * 32, 16, 1024);
              const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
              umma_bf16(acc_base + (uint32_t)(mt * BN), ad, bd, idesc, (first | k) != 0);
            }
        }
        umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
      }
      s += NI;
      if (s >= STAGES) { s -= STAGES; ph ^= 1u; }
    }
    umma_commit(tmem_full);           // one arrival per issuer: the barrier completes when every accumulator set is final
    if (who == 0) DBG_T(3);
  } else if (warp >= 2 && warp < 6) {
    // ===== epilogue: TMEM -> registers -> bias/activation -> bf16 NHWC =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    int row = q * 32 + lane;                // accumulator row = grid pixel inside the tile (first accumulator)
    int wl = row % P.wt, hl = (row / P.wt) % P.ht, nl = row / (P.wt * P.ht);
    int a = a0 + hl, b = b0 + wl, n = n0 + nl;
    int oy = a * P.ostride + P.oy0[cls], ox = b * P.ostride + P.ox0[cls];
    bool valid = n < P.N && oy < P.OH && ox < P.OW;
    __nv_bfloat16* out = P.y + ((long long)(n * P.OH + oy) * P.OW + ox) * P.ldy + n_col0;
    mbar_wait_warp(tmem_full, 0, lane);
    if (threadIdx.x == 64) DBG_T(4);
    tc_fence_after();
    if constexpr (BN == 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16), r);
      if (valid) {
        if (P.y32) {
#pragma unroll
          for (int co = 0; co < 16; ++co) {
            if (co < P.nout_real) {
              const float f = __uint_as_float(r[co]) + (P.bias ? P.bias[co] : 0.f);
              P.y32[(((long long)n * P.nout_real + co) * P.OH + oy) * P.OW + ox] = act_fwd(P.act, f);
            }
          }
        } else {   // 8-channel NHWC (gradient w.r.t. a packed thin input): one 16-byte store
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(out) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    } else if (P.part_out == nullptr) {
      constexpr int PITCH = BN * 2 + 16;
      if constexpr (EP) {
      // stage the bf16 tile in (now idle) pipeline smem, one row per thread, then store whole rows coalesced.  With a second
      // output (P.y2) the accumulators are read from TMEM a second time and go through the same staging rows.
      const float* bias = P.bias;
      const float* scale = P.ep_scale;
      const int npass = P.y2 ? 2 : 1;
#pragma unroll 1
      for (int op = 0; op < npass; ++op) {
      const float slope = act_slope(op == 0 ? P.act : P.act2);
      __nv_bfloat16* const ybase = op == 0 ? P.y : P.y2;
      const int ldo = op == 0 ? P.ldy : P.ldy2;
      if (op) __syncwarp();                     // this warp's row stores of the first pass have read the staging rows
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        {                                       // rows mt*128.. of the tile live in accumulator mt
          row = mt * TC_BM + q * 32 + lane;
          wl = row % P.wt; hl = (row / P.wt) % P.ht; nl = row / (P.wt * P.ht);
          a = a0 + hl; b = b0 + wl; n = n0 + nl;
          oy = a * P.ostride + P.oy0[cls]; ox = b * P.ostride + P.ox0[cls];
          valid = n < P.N && oy < P.OH && ox < P.OW;
          out = ybase + ((long long)(n * P.OH + oy) * P.OW + ox) * ldo + n_col0;
        }
        const uint32_t stg = smem_u32(smem) + (uint32_t)row * PITCH;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c0), r);
          if constexpr (NI == 2) {           // add the second issuer's accumulator
            uint32_t r2[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(MT * BN + mt * BN + c0), r2);
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
              if (scale) { f0 *= __ldg(scale + n_col0 + c0 + v * 8 + 2 * e); f1 *= __ldg(scale + n_col0 + c0 + v * 8 + 2 * e + 1); }
              if (bias) { f0 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e); f1 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e + 1); }
              f0 = act_piecewise(f0, slope); f1 = act_piecewise(f1, slope);
              const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              w[e] = valid ? *reinterpret_cast<const uint32_t*>(&h) : 0u;     // rows outside the tensor count as zeros below
            }
            st_shared_v4(stg + c0 * 2 + v * 16, w[0], w[1], w[2], w[3]);
          }
        }
        __syncwarp();
        if (threadIdx.x == 64) DBG_T(7);
        constexpr int LPR = BN * 2 / 16;     // lanes per output row (16-byte pieces)
        constexpr int RPI = 32 / LPR;        // rows per warp-wide store instruction
        const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(out) : 0ull;
        const uint32_t wbase = smem_u32(smem) + (uint32_t)(mt * TC_BM + q * 32) * PITCH;
#pragma unroll 4
        for (int i = 0; i < 32; i += RPI) {
          const int rr = i + lane / LPR;
          const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
          if (pr) {
            const uint4 v = ld_shared_v4(wbase + (uint32_t)rr * PITCH + (lane % LPR) * 16);
            *reinterpret_cast<uint4*>(pr + (lane % LPR) * 16) = v;
          }
        }
      }
      }
      } else {
      // stage the bf16 tile in (now idle) pipeline smem, one row per thread, then store whole rows coalesced
      const float slope = act_slope(P.act);
      const float* bias = P.bias;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        if (mt > 0) {                           // rows 128.. of the tile live in the second accumulator
          row = mt * TC_BM + q * 32 + lane;
          wl = row % P.wt; hl = (row / P.wt) % P.ht; nl = row / (P.wt * P.ht);
          a = a0 + hl; b = b0 + wl; n = n0 + nl;
          oy = a * P.ostride + P.oy0[cls]; ox = b * P.ostride + P.ox0[cls];
          valid = n < P.N && oy < P.OH && ox < P.OW;
          out = P.y + ((long long)(n * P.OH + oy) * P.OW + ox) * P.ldy + n_col0;
        }
        const uint32_t stg = smem_u32(smem) + (uint32_t)row * PITCH;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c0), r);
          if constexpr (NI == 2) {           // add the second issuer's accumulator
            uint32_t r2[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(MT * BN + mt * BN + c0), r2);
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
              if (bias) { f0 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e); f1 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e + 1); }
              f0 = act_piecewise(f0, slope); f1 = act_piecewise(f1, slope);
              const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              w[e] = valid ? *reinterpret_cast<const uint32_t*>(&h) : 0u;     // rows outside the tensor count as zeros below
            }
            st_shared_v4(stg + c0 * 2 + v * 16, w[0], w[1], w[2], w[3]);
          }
        }
        __syncwarp();
        if (threadIdx.x == 64) DBG_T(7);
        constexpr int LPR = BN * 2 / 16;     // lanes per output row (16-byte pieces)
        constexpr int RPI = 32 / LPR;        // rows per warp-wide store instruction
        const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(out) : 0ull;
        const uint32_t wbase = smem_u32(smem) + (uint32_t)(mt * TC_BM + q * 32) * PITCH;
#pragma unroll 4
        for (int i = 0; i < 32; i += RPI) {
          const int rr = i + lane / LPR;
          const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
          if (pr) {
            const uint4 v = ld_shared_v4(wbase + (uint32_t)rr * PITCH + (lane % LPR) * 16);
            *reinterpret_cast<uint4*>(pr + (lane % LPR) * 16) = v;
          }
        }
      }
      }
      if (P.bn_acc) {
        // BatchNorm statistics of this tile from the staged bf16 values: thread (rg, cg) sums 8 channels over its row
        // group, the row groups are combined through smem, one fp64 atomic pair per channel and CTA
        constexpr int CG = BN / 8, RG = 128 / CG, RPG = MT * TC_BM / RG;
        asm volatile("bar.sync 1, 128;" ::: "memory");              // all four epilogue warps have staged their rows
        const int et = threadIdx.x - 64, cg = et % CG, rg = et / CG;
        float s8[8], q8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s8[i] = 0.f; q8[i] = 0.f; }
        const uint32_t sbase = smem_u32(smem) + (uint32_t)(rg * RPG) * PITCH + cg * 16;
#pragma unroll 4
        for (int r = 0; r < RPG; ++r) {
          const uint4 u = ld_shared_v4(sbase + (uint32_t)r * PITCH);
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float fa = __uint_as_float(w4[i] << 16), fb = __uint_as_float(w4[i] & 0xffff0000u);
            s8[2 * i] += fa; q8[2 * i] = fmaf(fa, fa, q8[2 * i]);
            s8[2 * i + 1] += fb; q8[2 * i + 1] = fmaf(fb, fb, q8[2 * i + 1]);
          }
        }
        float* red = reinterpret_cast<float*>(smem + SM::STAT_OFFSET);      // [2][RG][BN]
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[rg * BN + cg * 8 + i] = s8[i]; red[RG * BN + rg * BN + cg * 8 + i] = q8[i]; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int slot = (int)((blockIdx.x + blockIdx.z) % STCGAN_BN_SLOTS);
        double* acc = P.bn_acc + (long long)slot * 2 * P.bn_c + n_col0;
        for (int c = et; c < BN; c += 128) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < RG; ++g2) { ts += red[g2 * BN + c]; tq += red[RG * BN + g2 * BN + c]; }
          atomicAdd(acc + c, (double)ts);
          atomicAdd(acc + P.bn_c + c, (double)tq);
        }
      }
    } else {
      // split-K partial: fp32 tile through smem, vector red.global.add of whole rows
      constexpr int PITCH = BN * 4 + 16;
      const uint32_t stg = smem_u32(smem) + (uint32_t)row * PITCH;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        if constexpr (NI == 2) {
          uint32_t r2[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(MT * BN + c0), r2);
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) st_shared_v4(stg + c0 * 4 + v * 16, r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
      }
      __syncwarp();
      constexpr int LPR = BN * 4 / 16;     // 32 (BN=128) or 16 (BN=64)
      constexpr int RPI = 32 / LPR;
      float* po = P.part_out + ((long long)(n * P.OH + oy) * P.OW + ox) * P.part_ld + n_col0;
      const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(po) : 0ull;
      const uint32_t wbase = smem_u32(smem) + (uint32_t)(q * 32) * PITCH;
#pragma unroll 4
      for (int i = 0; i < 32; i += RPI) {
        const int rr = i + lane / LPR;
        const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
        if (pr) red_add_v4(reinterpret_cast<float*>(pr) + (lane % LPR) * 4, ld_shared_v4(wbase + (uint32_t)rr * PITCH + (lane % LPR) * 16));
      }
    }
  }

  if (threadIdx.x == 64) DBG_T(5);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<SM::TMEM_COLS>(tmem_base);
  }
  if (threadIdx.x == 0) DBG_T(6);
}

// ---------------------------------------------------------------------------------------------
// persistent forward / dgrad kernel: one CTA per SM walks a static list of output tiles.  The smem ring runs straight
// through tile boundaries, the accumulator is double-buffered in TMEM (2 x BN columns), and the epilogue of tile i
// (TMEM -> registers -> bias/act -> bf16 -> staging smem -> coalesced row stores) overlaps the main loop of tile i+1.
// Removes the per-tile prologue (barrier init, TMEM alloc, first-TMA latency) and the exposed epilogue that the
// one-tile-per-CTA kernel above pays: measured per tile 0.25 + 1.4 + 2.2 us against main loops of 7-28 us.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct PersistSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PITCH = BN * 2 + 16;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGING_OFFSET + TC_BM * PITCH;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};


template <int BN, int STAGES, bool EP = false>
__global__ void __launch_bounds__(192, 1)
tapgemm_tc_persistent_kernel(const __grid_constant__ TapGemmParams P, int m_tiles, int n_tiles, int total_tiles) {
  pdl_trigger();
  using SM = PersistSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iters = P.ntaps * P.kchunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.bmap);
    tma_prefetch_desc(&P.amap[0]);
    tma_prefetch_desc(&P.amap[1]);
    tma_prefetch_desc(&P.amap[2]);
    tma_prefetch_desc(&P.amap[3]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is independent of the preceding kernel; global memory is touched only below

  if (threadIdx.x == 0) {
    // ===== TMA producer =====
    uint32_t g = 0;                                  // global k-iteration counter: the ring never restarts
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = t % m_tiles, nt = (t / m_tiles) % n_tiles, cls = t / (m_tiles * n_tiles);
      const int tw = mt % P.tiles_w, th = (mt / P.tiles_w) % P.tiles_h, tn = mt / (P.tiles_w * P.tiles_h);
      const int b0 = tw * P.wt, a0 = th * P.ht, n0 = tn * P.nt, n_col0 = nt * BN;
      for (int it = 0; it < iters; ++it, ++g) {
        const int s = g % STAGES;
        const uint32_t ph = (g / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int j = it / P.kchunks, kc = it % P.kchunks;
        uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
        uint8_t* b_dst = a_dst + SM::A_BYTES;
        mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
        if (P.thin_k) {
          tma_load_5d(&P.amap[0], &full_bar[s], a_dst, 0, b0, a0, n0, 2 * kc);
        } else {
          tma_load_4d(&P.amap[P.tview[cls][j]], &full_bar[s], a_dst, kc * TC_BK, b0 + P.tdx[cls][j], a0 + P.tdy[cls][j], n0);
        }
        tma_load_2d(&P.bmap, &full_bar[s], b_dst, kc * TC_BK, (int)P.twt[cls][j] * P.Nout + n_col0);
      }
    }
  } else if (threadIdx.x == 32) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(TC_BM, BN, 0, 0);
    uint32_t g = 0, li = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
      const uint32_t buf = li & 1u;
      mbar_wait(&tmem_empty[buf], ((li >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * BN;
      for (int it = 0; it < iters; ++it, ++g) {
        const int s = g % STAGES;
        const uint32_t ph = (g / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint32_t b_addr = a_addr + SM::A_BYTES;
        if (P.thin_k) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc(a_addr + (k >> 1) * 8192 + (k & 1) * 32, 16, 512, 4);
            const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_addr, ad, bd, idesc, (it | k) != 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_addr, ad, bd, idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full[buf]);
    }
  } else if (warp >= 2) {
    // ===== epilogue =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wl = row % P.wt, hl = (row / P.wt) % P.ht, nl = row / (P.wt * P.ht);
    const float slope = act_slope(P.act);
    const float* bias = P.bias;
    [[maybe_unused]] const float* scale = P.ep_scale;
    const uint32_t stg = smem_u32(smem + SM::STAGING_OFFSET) + (uint32_t)row * SM::PITCH;
    const uint32_t wbase = smem_u32(smem + SM::STAGING_OFFSET) + (uint32_t)(q * 32) * SM::PITCH;
    constexpr int LPR = BN * 2 / 16;
    constexpr int RPI = 32 / LPR;
    uint32_t li = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
      const int mt = t % m_tiles, nt = (t / m_tiles) % n_tiles, cls = t / (m_tiles * n_tiles);
      const int tw = mt % P.tiles_w, th = (mt / P.tiles_w) % P.tiles_h, tn = mt / (P.tiles_w * P.tiles_h);
      const int a = th * P.ht + hl, b = tw * P.wt + wl, n = tn * P.nt + nl, n_col0 = nt * BN;
      const int oy = a * P.ostride + P.oy0[cls], ox = b * P.ostride + P.ox0[cls];
      const bool valid = n < P.N && oy < P.OH && ox < P.OW;
      __nv_bfloat16* out = P.y + ((long long)(n * P.OH + oy) * P.OW + ox) * P.ldy + n_col0;
      const uint32_t buf = li & 1u;
      mbar_wait_warp(&tmem_full[buf], (li >> 1) & 1u, lane);
      tc_fence_after();
      if constexpr (EP) {
      const int npass = P.y2 ? 2 : 1;                 // second output: the accumulators are read twice (see TapGemmParams)
#pragma unroll 1
      for (int op = 0; op < npass; ++op) {
        const float slope_o = op == 0 ? slope : act_slope(P.act2);
        __nv_bfloat16* const out_o = op == 0 ? out : P.y2 + ((long long)(n * P.OH + oy) * P.OW + ox) * P.ldy2 + n_col0;
        __syncwarp();                                 // previous row stores of this warp have read the staging rows
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c0, r);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
              if (scale) { f0 *= __ldg(scale + n_col0 + c0 + v * 8 + 2 * e); f1 *= __ldg(scale + n_col0 + c0 + v * 8 + 2 * e + 1); }
              if (bias) { f0 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e); f1 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e + 1); }
              f0 = act_piecewise(f0, slope_o); f1 = act_piecewise(f1, slope_o);
              const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              w[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            st_shared_v4(stg + c0 * 2 + v * 16, w[0], w[1], w[2], w[3]);
          }
        }
        if (op == npass - 1) {
          // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld): hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        } else {
          __syncwarp();
        }
        const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(out_o) : 0ull;
#pragma unroll 4
        for (int i = 0; i < 32; i += RPI) {
          const int rr = i + lane / LPR;
          const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
          if (pr) {
            const uint4 v = ld_shared_v4(wbase + (uint32_t)rr * SM::PITCH + (lane % LPR) * 16);
            *reinterpret_cast<uint4*>(pr + (lane % LPR) * 16) = v;
          }
        }
      }
      } else {
      __syncwarp();                                   // previous tile's row stores of this warp have read the staging rows
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c0, r);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
            if (bias) { f0 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e); f1 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e + 1); }
            f0 = act_piecewise(f0, slope); f1 = act_piecewise(f1, slope);
            const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          st_shared_v4(stg + c0 * 2 + v * 16, w[0], w[1], w[2], w[3]);
        }
      }
      // all TMEM reads of this buffer are complete (tcgen05.wait::ld inside tmem_ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(out) : 0ull;
#pragma unroll 4
      for (int i = 0; i < 32; i += RPI) {
        const int rr = i + lane / LPR;
        const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
        if (pr) {
          const uint4 v = ld_shared_v4(wbase + (uint32_t)rr * SM::PITCH + (lane % LPR) * 16);
          *reinterpret_cast<uint4*>(pr + (lane % LPR) * 16) = v;
        }
      }
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

// ---------------------------------------------------------------------------------------------
// CTA-pair forward / dgrad kernel (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN2 tile -- each CTA
// owns 128 output pixels (its own A tile and its own 128 accumulator lanes in its own TMEM) and stages HALF of the
// BN2-wide weight tile; the leader CTA's MMA thread issues tcgen05.mma.cta_group::2, which reads both halves.  Per SM
// this halves the weight bytes that must be pulled through L2/TMA per FLOP (the measured limiter of the single-CTA
// kernel: ~80 B/cycle/SM of operand ingest against the 128 B/cycle a 128x128 tile needs) and the smem read traffic.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's mbarrier (executed by both CTAs of the pair)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint32_t leader_bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once all prior MMAs of the pair have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

template <int BN2, int STAGES>
struct PairSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;             // this CTA's 128 pixels
  static constexpr int B_BYTES = (BN2 / 2) * TC_BK * 2;         // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PITCH = BN2 * 2 + 16;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int STAGING = TC_BM * PITCH;
  static constexpr int BAR_OFFSET = RING > STAGING ? RING : STAGING;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN2, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192)
tapgemm_tc_pair_kernel(const __grid_constant__ TapGemmParams P) {
  pdl_trigger();
  using SM = PairSmem<BN2, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cls = blockIdx.z;
  const int n_col0 = blockIdx.y * BN2;
  const int mt = blockIdx.x;                      // the pair (2i, 2i+1) shares blockIdx.y / z
  const int tw = mt % P.tiles_w;
  const int th = (mt / P.tiles_w) % P.tiles_h;
  const int tn = mt / (P.tiles_w * P.tiles_h);    // may exceed the tensor for the padding CTA of an odd tile count: all OOB
  const int b0 = tw * P.wt, a0 = th * P.ht, n0 = tn * P.nt;
  const int iters = P.ntaps * P.kchunks;

  if (threadIdx.x == 0) {
    DBG_T(0);
    tma_prefetch_desc(&P.bmap);
    tma_prefetch_desc(&P.amap[0]);
    tma_prefetch_desc(&P.amap[1]);
    tma_prefetch_desc(&P.amap[2]);
    tma_prefetch_desc(&P.amap[3]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<BN2>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                              // barriers of both CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is independent of the preceding kernel; global memory is touched only below
  if (threadIdx.x == 0) DBG_T(1);

  if (threadIdx.x == 0) {
    // ===== TMA producer (both CTAs): own A tile + own half of B; bytes are credited to the leader's full barrier =====
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int j = it / P.kchunks, kc = it % P.kchunks;
      uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
      uint8_t* b_dst = a_dst + SM::A_BYTES;
      const uint32_t lbar = smem_u32(&full_bar[s]) & PEER_BIT_MASK;
      if (leader) mbar_expect_tx(&full_bar[s], 2 * SM::STAGE_BYTES);
      else mbar_arrive_cluster(lbar);
      tma_load_4d_pair(&P.amap[P.tview[cls][j]], lbar, a_dst, kc * TC_BK, b0 + P.tdx[cls][j], a0 + P.tdy[cls][j], n0);
      tma_load_2d_pair(&P.bmap, lbar, b_dst, kc * TC_BK, (int)P.twt[cls][j] * P.Nout + n_col0 + (int)rank * (BN2 / 2));
    }
  } else if (threadIdx.x == 32 && leader) {
    // ===== MMA issuer (leader CTA only): 256 x BN2 x 16 per instruction across the pair =====
    constexpr uint32_t idesc = make_idesc(256, BN2, 0, 0);
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      if (it == 0) DBG_T(2);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
      const uint32_t b_addr = a_addr + SM::A_BYTES;
#pragma unroll
      for (int k = 0; k < TC_BK / 16; ++k) {
        const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);
        const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
        umma_bf16_pair(tmem_base, ad, bd, idesc, (it | k) != 0);
      }
      umma_commit_pair(&empty_bar[s]);             // frees this stage in BOTH CTAs
    }
    umma_commit_pair(tmem_full);
    DBG_T(3);
  } else if (warp >= 2) {
    // ===== epilogue (both CTAs): own 128 rows x BN2 columns =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wl = row % P.wt, hl = (row / P.wt) % P.ht, nl = row / (P.wt * P.ht);
    const int a = a0 + hl, b = b0 + wl, n = n0 + nl;
    const int oy = a * P.ostride + P.oy0[cls], ox = b * P.ostride + P.ox0[cls];
    const bool valid = n < P.N && oy < P.OH && ox < P.OW;
    __nv_bfloat16* out = P.y + ((long long)(n * P.OH + oy) * P.OW + ox) * P.ldy + n_col0;
    mbar_wait_warp(tmem_full, 0, lane);
    if (threadIdx.x == 64) DBG_T(4);
    tc_fence_after();
    const uint32_t stg = smem_u32(smem) + (uint32_t)row * SM::PITCH;
    const float slope = act_slope(P.act);
    const float* bias = P.bias;
#pragma unroll 1
    for (int c0 = 0; c0 < BN2; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f0 = __uint_as_float(r[v * 8 + 2 * e]), f1 = __uint_as_float(r[v * 8 + 2 * e + 1]);
          if (bias) { f0 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e); f1 += __ldg(bias + n_col0 + c0 + v * 8 + 2 * e + 1); }
          f0 = act_piecewise(f0, slope); f1 = act_piecewise(f1, slope);
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        st_shared_v4(stg + c0 * 2 + v * 16, w[0], w[1], w[2], w[3]);
      }
    }
    __syncwarp();
    constexpr int LPR = BN2 * 2 / 16;    // 32 (BN2 = 256: one row per instruction) or 16
    constexpr int RPI = 32 / LPR;
    const unsigned long long myp = valid ? reinterpret_cast<unsigned long long>(out) : 0ull;
    const uint32_t wbase = smem_u32(smem) + (uint32_t)(q * 32) * SM::PITCH;
#pragma unroll 4
    for (int i = 0; i < 32; i += RPI) {
      const int rr = i + lane / LPR;
      const unsigned long long pr = __shfl_sync(0xffffffffu, myp, rr);
      if (pr) {
        const uint4 v = ld_shared_v4(wbase + (uint32_t)rr * SM::PITCH + (lane % LPR) * 16);
        *reinterpret_cast<uint4*>(pr + (lane % LPR) * 16) = v;
      }
    }
  }

  if (threadIdx.x == 64) DBG_T(5);
  tc_fence_before();
  cluster_sync_all();                              // the peer may still be reading this CTA's smem / arriving on its barriers
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair<BN2>(tmem_base);
  }
  if (threadIdx.x == 0) DBG_T(6);
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel: G[tap][d0 tile 128][d1 tile BN] += sum over this CTA's pixel tiles
// ---------------------------------------------------------------------------------------------
struct alignas(64) WgradParams {
  CUtensorMap smap;        // S  [N, SH, SW, D0]
  CUtensorMap lmap[4];     // L views (parity views for stride 2)
  int8_t tdy[16], tdx[16], tview[16];
  int wt, ht, nt;          // pixel tile box: wt*ht*nt = 64
  int tiles_w, tiles_h, tiles_n;
  int tiles_per_split;     // pixel tiles handled by one CTA (blockIdx.z = split)
  int D0, D1;
  float* G;
  // thin mode: D[(tap,c)][d] = sum_q Twin[q,(tap,c)] * F[q,d] with T a zero-bordered 8-channel tensor (5D im2col map
  // in lmap[0]) and F the fat tensor (smap).  M = 128 im2col columns, N = BN channels of F.
  int thin;
  int thin_c;              // real channels of T (<= 8)
  int fat_is_dim0;         // F's channels index the weight's dim0 (else dim1)
  int flip;                // window taps are the flipped kernel taps (stride-1 dgrad-style anchoring)
  int Dfat;
};

template <int BN, int STAGES>
struct WgradSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;        // two [64 px][64 ch] blocks
  static constexpr int B_BYTES = BN * 64 * 2;         // BN/64 blocks
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192)
tapwgrad_tc_kernel(const __grid_constant__ WgradParams P) {
  pdl_trigger();
  using SM = WgradSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles1 = P.D1 / BN;
  const int d0_0 = (blockIdx.x / tiles1) * 128, d1_0 = (blockIdx.x % tiles1) * BN;
  const int tap = blockIdx.y;
  const int total_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  const int t_begin = blockIdx.z * P.tiles_per_split;
  int t_end = t_begin + P.tiles_per_split;
  if (t_end > total_tiles) t_end = total_tiles;
  const int iters = t_end - t_begin;     // host guarantees >= 1

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.smap);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is independent of the preceding kernel; global memory is touched only below

  if (threadIdx.x == 0) {
    const CUtensorMap* lm = &P.lmap[P.tview[tap]];
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      const int t = t_begin + it;
      const int b0 = (t % P.tiles_w) * P.wt, a0 = ((t / P.tiles_w) % P.tiles_h) * P.ht, n0 = (t / (P.tiles_w * P.tiles_h)) * P.nt;
      uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
      uint8_t* b_dst = a_dst + SM::A_BYTES;
      mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
      // two TMA instructions per stage: every operand tile (2 or BN/64 channel blocks, or the 4 kernel rows of the thin
      // im2col view) is one 5-D box whose slowest dimension enumerates the [64 px][row] blocks
      if (P.thin) {
        tma_load_5d(&P.lmap[0], &full_bar[s], a_dst, 0, b0, a0, n0, 0);                       // M rows (kh, kw, c)
        tma_load_5d(&P.smap, &full_bar[s], b_dst, 0, b0, a0, n0, d1_0 / 64);                  // N = BN channels of F
      } else {
        tma_load_5d(&P.smap, &full_bar[s], a_dst, 0, b0, a0, n0, d0_0 / 64);                  // M = 128 channels of S
        tma_load_5d(lm, &full_bar[s], b_dst, 0, b0 + P.tdx[tap], a0 + P.tdy[tap], n0, d1_0 / 64);   // N = BN channels of L
      }
    }
  } else if (threadIdx.x == 32) {
    constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);     // both operands MN-major
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
      const uint32_t b_addr = a_addr + SM::A_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // 64 pixels per stage = 4 x K16; 16 K-rows = 2048 B (1024 B in 64-byte-row blocks)
        const uint64_t ad = P.thin ? make_smem_desc(a_addr + k * 1024, 4096, 512, 4)
                                   : make_smem_desc(a_addr + k * 2048, 8192, 1024);
        const uint64_t bd = make_smem_desc(b_addr + k * 2048, 8192, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, (it | k) != 0);
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(tmem_full);
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait_warp(tmem_full, 0, lane);
    tc_fence_after();
    if (P.thin) {
      // stage the fp32 tile [(tap, c)][d] in (now idle) pipeline smem, then accumulate with consecutive threads on
      // consecutive addresses of G: for one tap the (d, c) block -- or the (c, d) rows -- of this CTA is contiguous
      // (scalar atomics issued row-per-thread hit 32 different cache lines per instruction and serialised the 296 CTAs)
      constexpr int PITCH = BN + 1;
      float* stage = reinterpret_cast<float*>(smem);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
        for (int e = 0; e < 32; ++e) stage[row * PITCH + c0 + e] = __uint_as_float(r[e]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = threadIdx.x - 64, tc = P.thin_c, per_tap = BN * tc;
      for (int t = 0; t < 16; ++t) {
        const int wtap = P.flip ? 15 - t : t;
        for (int j = et; j < per_tap; j += 128) {
          if (P.fat_is_dim0) {           // G[wtap][d][c]: j = d_local * thin_c + c
            const int dl = j / tc, c = j - dl * tc;
            atomicAdd(P.G + ((long long)wtap * P.Dfat + d1_0) * tc + j, stage[(t * 8 + c) * PITCH + dl]);
          } else {                       // G[wtap][c][d]: j = c * BN + d_local
            const int c = j / BN, dl = j - c * BN;
            atomicAdd(P.G + ((long long)wtap * tc + c) * P.Dfat + d1_0 + dl, stage[(t * 8 + c) * PITCH + dl]);
          }
        }
      }
    } else {
      // fp32 tile through (now idle) pipeline smem, then one vector red.global.add per 16 bytes of a whole row:
      // a warp-wide instruction touches one or two contiguous rows instead of 32 different cache lines
      constexpr int PITCH = BN * 4 + 16;
      const uint32_t stg = smem_u32(smem) + (uint32_t)row * PITCH;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
        for (int v = 0; v < 8; ++v) st_shared_v4(stg + c0 * 4 + v * 16, r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
      }
      __syncwarp();
      constexpr int LPR = BN * 4 / 16;
      constexpr int RPI = 32 / LPR;
      float* out = P.G + ((long long)tap * P.D0 + d0_0 + q * 32) * P.D1 + d1_0;
      const uint32_t wbase = smem_u32(smem) + (uint32_t)(q * 32) * PITCH;
#pragma unroll 4
      for (int i = 0; i < 32; i += RPI) {
        const int rr = i + lane / LPR;
        red_add_v4(out + (long long)rr * P.D1 + (lane % LPR) * 4, ld_shared_v4(wbase + (uint32_t)rr * PITCH + (lane % LPR) * 16));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// choose a box (wt, ht, nt) of `total` grid pixels (powers of two) that wastes the fewest MMA rows
static void choose_tile(int total, int GW, int GH, int N, int* wt, int* ht, int* nt) {
  double best = -1.0;
  for (int w = 1; w <= total; w *= 2)
    for (int h = 1; w * h <= total; h *= 2) {
      const int n = total / (w * h);
      const long long tiles = (long long)((GW + w - 1) / w) * ((GH + h - 1) / h) * ((N + n - 1) / n);
      const double util = (double)GW * GH * N / ((double)tiles * total);
      // prefer wide boxes on ties (longer contiguous runs for TMA)
      const double score = util + 1e-6 * w + 1e-9 * h;
      if (score > best) { best = score; *wt = w; *ht = h; *nt = n; }
    }
}


// profiling aid: with STCGAN_TC_DEBUG_TIMES=<path> every tapgemm launch dumps per-CTA phase timestamps (synchronous!)
static int dump_debug_times(const TapGemmParams& P0, dim3 grid, cudaStream_t st, int bn,
                            int (*launch)(const TapGemmParams&, dim3, cudaStream_t)) {
  const char* path = getenv("STCGAN_TC_DEBUG_TIMES");
  if (!path) return launch(P0, grid, st);
  TapGemmParams P = P0;
  const size_t n = (size_t)grid.x * grid.y * grid.z * 8;
  long long* d = nullptr;
  cudaMalloc(&d, n * sizeof(long long));
  cudaMemset(d, 0, n * sizeof(long long));
  P.dbg = d;
  int rc = launch(P, grid, st);
  cudaStreamSynchronize(st);
  long long* h = (long long*)malloc(n * sizeof(long long));
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  FILE* f = fopen(path, "a");
  if (f) {
    fprintf(f, "launch bn=%d grid=%u,%u,%u iters=%d thin_k=%d ksplit=%d\n", bn, grid.x, grid.y, grid.z, P.ntaps * P.kchunks, P.thin_k, P.ksplit);
    for (size_t c = 0; c < n / 8; ++c) {
      for (int k = 0; k < 8; ++k) fprintf(f, "%lld ", h[c * 8 + k]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
  free(h); cudaFree(d);
  return rc;
}

template <int BN, int STAGES, int MT = 1, int NI = 1>
static int launch_tapgemm_raw(const TapGemmParams& P, dim3 grid, cudaStream_t st);

template <int BN, int STAGES, int MT = 1, int NI = 1>
static int launch_tapgemm(const TapGemmParams& P, dim3 grid, cudaStream_t st) {
  return dump_debug_times(P, grid, st, BN, &launch_tapgemm_raw<BN, STAGES, MT, NI>);
}

template <int BN, int STAGES, int MT, int NI>
static int launch_tapgemm_raw(const TapGemmParams& P, dim3 grid, cudaStream_t st) {
  using SM = TapGemmSmem<BN, STAGES, MT, NI>;
  if constexpr (NI == 1 && BN >= 64) {
    if (P.ep_scale || P.y2) {          // extended epilogue: its own instantiation (see tapgemm_tc_kernel)
      static bool configured_ep = false;
      if (!configured_ep) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_kernel<BN, STAGES, MT, NI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
        if (e != cudaSuccess) return (int)e;
        configured_ep = true;
      }
      launch_k(tapgemm_tc_kernel<BN, STAGES, MT, NI, true>, grid, 192, SM::TOTAL, st, P);
      return finish_launch();
    }
  }
  if (P.ep_scale || P.y2) return STCGAN_EUNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_kernel<BN, STAGES, MT, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(tapgemm_tc_kernel<BN, STAGES, MT, NI>, grid, NI == 2 ? 224 : 192, SM::TOTAL, st, P);
  return finish_launch();
}

template <int BN, int STAGES>
static int launch_tapgemm_persistent(const TapGemmParams& P, int m_tiles, int n_tiles, int nclass, cudaStream_t st,
                                     int ctas_per_sm = 1) {
  using SM = PersistSmem<BN, STAGES>;
  const int total = m_tiles * n_tiles * nclass;
  const int cap = 148 * ctas_per_sm;
  const int grid = total < cap ? total : cap;
  if (P.ep_scale || P.y2) {            // extended epilogue: its own instantiation (see tapgemm_tc_kernel)
    static bool configured_ep = false;
    if (!configured_ep) {
      cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_persistent_kernel<BN, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
      if (e != cudaSuccess) return (int)e;
      configured_ep = true;
    }
    launch_k(tapgemm_tc_persistent_kernel<BN, STAGES, true>, grid, 192, SM::TOTAL, st, P, m_tiles, n_tiles, total);
    return finish_launch();
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_persistent_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(tapgemm_tc_persistent_kernel<BN, STAGES>, grid, 192, SM::TOTAL, st, P, m_tiles, n_tiles, total);
  return finish_launch();
}

// STCGAN_TC_PERSISTENT: 0 = never, 1 = every full-width forward/dgrad launch, unset = thin-K layers only (measured on
// B200: the operand stream of a 128-wide tile is bound by ~80 B/cycle/SM of TMA/L2 ingest, where 2 CTAs x 3 stages per SM
// beat 1 persistent CTA x 5 stages; the 2048-tile / 2-iteration thin-K layers gain 20 % from persistence)
template <int BN2, int STAGES>
static int launch_tapgemm_pair_raw(const TapGemmParams& P, dim3 grid, cudaStream_t st) {
  using SM = PairSmem<BN2, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_pair_kernel<BN2, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(tapgemm_tc_pair_kernel<BN2, STAGES>, grid, 192, SM::TOTAL, st, P);
  return finish_launch();
}

template <int BN2, int STAGES>
static int launch_tapgemm_pair(const TapGemmParams& P, int m_tiles, int n_tiles, int nclass, cudaStream_t st) {
  dim3 grid((unsigned)((m_tiles + 1) / 2 * 2), (unsigned)n_tiles, (unsigned)nclass);
  return dump_debug_times(P, grid, st, -BN2, &launch_tapgemm_pair_raw<BN2, STAGES>);
}

// STCGAN_TC_PAIR: 1 = use the CTA-pair (cta_group::2) kernel for full-width forward/dgrad launches with Nout % 256 == 0.
// Opt-in: it is parity-green (all GPU tests pass with it) but measured SLOWER than the single-CTA kernel on B200 so far:
// the leader's main loop runs at 0.9 us per 256x256x64 step against 0.42 us per 128x128x64 step of the single-CTA kernel
// (profiles/r01_tapgemm_cta_phase_times.txt); the operand stream per CTA, not the tensor pipe, is the limiter in both.
static int pair_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STCGAN_TC_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

// accumulator width of the single-CTA kernel.  STCGAN_TC_BN256=1 selects 128x256 tiles (2 stages, 2 CTAs/SM) where the layer
// allows it; measured on B200 it is parity-green but 12 % slower on the conv launches of a train step than 128x128 tiles
// with 3 stages, so the default stays 128.
static int bn_select(int Nout) {
  int wide;
  { const char* e = getenv("STCGAN_TC_BN256"); wide = (e && e[0] == '1') ? 1 : 0; }
  if (wide && Nout % 256 == 0) return 256;
  return Nout % 128 == 0 ? 128 : 64;
}

static int wgrad_split_floor() {  // STCGAN_WGRAD_SPLIT_CEIL=1 restores the old ceil() split count (experiments)
  const char* e = getenv("STCGAN_WGRAD_SPLIT_CEIL");
  return !(e && e[0] == '1');
}

static int mt_mode() {            // STCGAN_TC_MT: unset = automatic, 1 = one accumulator per CTA always, 2 = two whenever possible
  const char* e = getenv("STCGAN_TC_MT");
  return !e ? 0 : (e[0] == '2' ? 2 : 1);
}

static int dual_issue_mode() {
  const char* e = getenv("STCGAN_TC_DUAL");
  return (e && e[0] == '1') ? 1 : 0;
}

static int bn256_auto() {         // STCGAN_TC_BN256_AUTO=0 disables the wave-aware choice of 128x256 tiles
  const char* e = getenv("STCGAN_TC_BN256_AUTO");
  return !(e && e[0] == '0');
}

static int deep_ring_mode() {     // STCGAN_TC_DEEP=0 disables the 6/8-stage variants for single-wave launches
  static int v = -1;
  if (v < 0) { const char* e = getenv("STCGAN_TC_DEEP"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

static int bn256_stages() {     // STCGAN_TC_BN256_STAGES=2: two CTAs per SM with two stages each; default 4 stages, one CTA
  const char* e = getenv("STCGAN_TC_BN256_STAGES");
  return (e && e[0] == '2') ? 2 : 4;
}

static int persistent_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STCGAN_TC_PERSISTENT"); v = !e ? 2 : (e[0] == '0' ? 0 : 1); }
  return v;
}

// thin_n != 0: Nout <= 16 real output channels, wp packed with 16 (zero-padded) rows per tap; output goes to y32
// (NCHW fp32, bias + any activation) or, if y32 == NULL, to the first 8 channels of an NHWC bf16 tensor.
// fp32 split-K partial sums [P][Nout] -> bf16 y (pitch ldy) with bias + activation
struct FinishExtra {          // see EpilogueExtra; OW/OH = true output extent of the partial sums, HC/WC = store crop
  const float* scale; __nv_bfloat16* y2; int ldy2, act2; int OH, OW, HC, WC;
};

__global__ void __launch_bounds__(256)
splitk_finish_kernel(float* __restrict__ part, long long P, int Nout, const float* __restrict__ bias, int act,
                     __nv_bfloat16* __restrict__ y, int ldy, double* __restrict__ bn_acc, int clear, const FinishExtra ex) {
  pdl_prologue();
  const int quads = Nout / 4;
  const float slope = act_slope(act);
  if (bn_acc == nullptr) {
    const float slope2 = act_slope(ex.act2);
    const bool crop = ex.HC != ex.OH || ex.WC != ex.OW;
    const long long total = P * quads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long p = i / quads; const int c = (int)(i % quads) * 4;
      const float4 v = *reinterpret_cast<const float4*>(part + p * Nout + c);
      if (clear) *reinterpret_cast<float4*>(part + p * Nout + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      long long pd = p;
      if (crop) {                                   // destination pixel inside the cropped extent (or skipped)
        const int ox = (int)(p % ex.OW); const long long t = p / ex.OW;
        const int oy = (int)(t % ex.OH); const long long n = t / ex.OH;
        if (oy >= ex.HC || ox >= ex.WC) continue;
        pd = (n * ex.HC + oy) * ex.WC + ox;
      }
      float f[4] = {v.x, v.y, v.z, v.w}, g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float z = f[e] * (ex.scale ? ex.scale[c + e] : 1.f) + (bias ? bias[c + e] : 0.f);
        f[e] = act_piecewise(z, slope); g[e] = act_piecewise(z, slope2);
      }
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
      uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(y + pd * ldy + c) = o;
      if (ex.y2) {
        const __nv_bfloat162 k0 = __floats2bfloat162_rn(g[0], g[1]), k1 = __floats2bfloat162_rn(g[2], g[3]);
        uint2 o2; o2.x = *reinterpret_cast<const uint32_t*>(&k0); o2.y = *reinterpret_cast<const uint32_t*>(&k1);
        *reinterpret_cast<uint2*>(ex.y2 + pd * ex.ldy2 + c) = o2;
      }
    }
    return;
  }
  // with BatchNorm statistics (host guarantees quads <= 256 and 256 % quads == 0): a thread keeps one channel quad and
  // walks down the rows (four independent row loads in flight); per-thread fp32 sums of the bf16-rounded outputs are combined
  // across the block's row groups in shared memory, then ONE fp64 atomic pair per channel and block goes into slot
  // blockIdx.x % SLOTS (round 1 issued them per thread: 262 k contended fp64 atomics made this kernel slower than the GEMM)
  __shared__ float red[2 * 256 * 4];
  const int qd = threadIdx.x % quads, rr = threadIdx.x / quads, rpb = 256 / quads, c = qd * 4;
  float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * rpb;
  for (long long p0 = (long long)blockIdx.x * rpb + rr; p0 < P; p0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long p = p0 + u * stride;
      v[u] = p < P ? *reinterpret_cast<const float4*>(part + p * Nout + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long p = p0 + u * stride;
      if (p >= P) continue;
      if (clear) *reinterpret_cast<float4*>(part + p * Nout + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) f[e] = act_piecewise(f[e] + (bias ? bias[c + e] : 0.f), slope);
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
      uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(y + p * ldy + c) = o;
      const float r[4] = {__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1)};
#pragma unroll
      for (int e = 0; e < 4; ++e) { s4[e] += r[e]; q4[e] = fmaf(r[e], r[e], q4[e]); }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { red[rr * Nout + c + e] = s4[e]; red[1024 + rr * Nout + c + e] = q4[e]; }
  __syncthreads();
  double* acc = bn_acc + (long long)(blockIdx.x % STCGAN_BN_SLOTS) * 2 * Nout;
  for (int j = threadIdx.x; j < Nout; j += 256) {
    float ts = 0.f, tq = 0.f;
    for (int g2 = 0; g2 < rpb; ++g2) { ts += red[g2 * Nout + j]; tq += red[1024 + g2 * Nout + j]; }
    atomicAdd(acc + j, (double)ts);
    atomicAdd(acc + Nout + j, (double)tq);
  }
}

int tapconv_tc(int geom_kind, const Geom& g, const void* x, int K, int ldx, const void* wp, const float* bias, int act,
               void* y, int Nout, int ldy, cudaStream_t st, int thin_n = 0, float* y32 = nullptr,
               float* ws = nullptr, long long ws_bytes = 0, double* bn_acc = nullptr, const EpilogueExtra* ex = nullptr) {
  if (bn_acc && (thin_n || act != STCGAN_ACT_NONE || Nout % 64 != 0)) return STCGAN_EUNSUPPORTED;
  if (ex && (bn_acc || thin_n)) return STCGAN_EUNSUPPORTED;
  if (ex && ex->y2 && (ex->ldy2 % 8 != 0 || !al16(ex->y2) || ex->act2 == STCGAN_ACT_TANH || ex->act2 == STCGAN_ACT_SIGMOID))
    return STCGAN_EUNSUPPORTED;
  // ws_bytes < 0: the workspace (|ws_bytes| bytes) is all zeros on entry and is left all zeros on exit (the finisher clears
  // what it reads): a persistent workspace saves the memset launch in front of every split-K convolution
  const bool ws_clean = ws_bytes < 0;
  if (ws_clean) ws_bytes = -ws_bytes;
  if (K % 64 != 0 || ldx % 8 != 0 || !al16(x) || !al16(wp)) return STCGAN_EUNSUPPORTED;
  if (!thin_n) {
    if (Nout % 64 != 0 || ldy % 8 != 0 || !al16(y)) return STCGAN_EUNSUPPORTED;
    if (act == STCGAN_ACT_TANH || act == STCGAN_ACT_SIGMOID) return STCGAN_EUNSUPPORTED;
  } else {
    if (Nout > 16 || Nout < 1) return STCGAN_EUNSUPPORTED;
    if (!y32 && (Nout > 8 || ldy % 8 != 0 || !al16(y) || bias || act != STCGAN_ACT_NONE)) return STCGAN_EUNSUPPORTED;
  }
  TapGemmParams P;
  memset(&P, 0, sizeof(P));
  P.nout_real = Nout; P.y32 = y32;
  P.bn_acc = bn_acc; P.bn_c = Nout;
  { const char* e = getenv("STCGAN_TC_DBGMODE"); P.dbg_mode = e ? atoi(e) : 0; }
  P.wp_raw = wp;
  const int n_rows = thin_n ? 16 : Nout;     // rows per tap in the packed weight matrix
  int GH = 0, GW = 0;
  for (int c = 0; c < g.nclass; ++c) {
    if (g.grid_h(c) > GH) GH = g.grid_h(c);
    if (g.grid_w(c) > GW) GW = g.grid_w(c);
    P.oy0[c] = g.oy0[c]; P.ox0[c] = g.ox0[c];
  }
  if (GH <= 0 || GW <= 0) return 0;
  choose_tile(TC_BM, GW, GH, g.N, &P.wt, &P.ht, &P.nt);
  P.tiles_w = (GW + P.wt - 1) / P.wt; P.tiles_h = (GH + P.ht - 1) / P.ht;
  // M tile of 2 x 128 pixels (two accumulators per CTA) for launches that would otherwise need more than one wave of
  // resident CTAs: half the CTAs, every weight tile feeds two MMAs, the per-stage wait + commit is paid once per 8 MMAs
  int MTsel = 1;
  if (!thin_n && Nout % 64 == 0) {
    const int bn1 = Nout % 128 == 0 ? 128 : 64;
    const long long ctas1 = (long long)P.tiles_w * P.tiles_h * ((g.N + P.nt - 1) / P.nt) * (Nout / bn1) * g.nclass;
    const int mode = mt_mode();
    // measured (tools/conv_probe.py, us per launch, one vs two accumulators): e2 fwd 23.5 -> 21.4, c3 dgrad 23.4 -> 21.7,
    // d3 fwd 35.8 -> 34.3; 128x64 tiles lose (c2 dgrad 34.6 -> 40.4, d2 fwd 54.7 -> 61.2: they live on residency, 3-4 CTAs
    // per SM), and where 128x256 tiles bring the launch into one wave those win instead (c4 fwd 51.1 vs 59.1)
    const long long c256 = Nout % 256 == 0 ? ctas1 / 2 : 0;
    const bool wide_wins = bn256_auto() && c256 >= 200 && c256 <= 296;
    if ((mode == 2 && ctas1 > 148) || (mode == 0 && bn1 == 128 && ctas1 > 296 && !wide_wins)) MTsel = 2;
  }
  if (MTsel == 2) {
    choose_tile(2 * TC_BM, GW, GH, g.N, &P.wt, &P.ht, &P.nt);
    P.tiles_w = (GW + P.wt - 1) / P.wt; P.tiles_h = (GH + P.ht - 1) / P.ht;
  }
  const int tiles_n = (g.N + P.nt - 1) / P.nt;
  P.GH = GH; P.GW = GW; P.N = g.N; P.OH = g.OH; P.OW = g.OW; P.ostride = g.ostride;
  P.ntaps = g.ntaps; P.kchunks = K / 64;
  P.Nout = Nout; P.ldy = ldy; P.act = act; P.bias = bias; P.y = static_cast<__nv_bfloat16*>(y);
  FinishExtra fx;
  fx.scale = nullptr; fx.y2 = nullptr; fx.ldy2 = 0; fx.act2 = 0; fx.OH = g.OH; fx.OW = g.OW; fx.HC = g.OH; fx.WC = g.OW;
  if (ex) {
    P.ep_scale = ex->scale; P.y2 = static_cast<__nv_bfloat16*>(ex->y2); P.ldy2 = ex->ldy2; P.act2 = ex->act2;
    fx.scale = ex->scale; fx.y2 = P.y2; fx.ldy2 = ex->ldy2; fx.act2 = ex->act2;
    if (ex->HC > 0 && ex->WC > 0) {
      if (ex->HC > g.OH || ex->WC > g.OW) return STCGAN_EINVAL;
      P.OH = ex->HC; P.OW = ex->WC;           // store extent (the tile grid below still covers the layer's true output)
      fx.HC = ex->HC; fx.WC = ex->WC;
    }
  }

  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const long long sn = (long long)g.IH * g.IW * ldx;
  int rc;
  if (g.istride == 1) {
    rc = encode_nhwc(&P.amap[0], xb, K, g.IW, g.IH, g.N, ldx, (long long)g.IW * ldx, sn, P.wt, P.ht, P.nt);
    if (rc) return rc;
    for (int v = 1; v < 4; ++v) P.amap[v] = P.amap[0];
  } else {
    for (int p = 0; p < 2; ++p)
      for (int q = 0; q < 2; ++q) {
        const long long vw = (g.IW - q + 1) / 2, vh = (g.IH - p + 1) / 2;   // rows 2a'+p < IH
        // a view can be empty when IH or IW == 1; keep a valid 1-wide map (never selected inside range).  Its base must
        // still be a mapped address: base + (p*IW + q) pixels can lie past the end of the tensor -- and of the allocation
        // segment -- and the TMA unit faults on an unmapped tensor-map base even if every coordinate is out of range
        // (observed: illegal address with a 2 x 1 x 1 x 512 input that ended on a 2 MB segment boundary)
        const bool empty = vw < 1 || vh < 1;
        rc = encode_nhwc(&P.amap[p * 2 + q], empty ? xb : xb + ((long long)p * g.IW + q) * ldx, K, vw, vh, g.N,
                         2LL * ldx, 2LL * g.IW * ldx, sn, P.wt, P.ht, P.nt);
        if (rc) return rc;
      }
  }
  // deep-K layers with few output tiles (the U-Net bottleneck) split their taps over extra CTAs and reduce in fp32; that
  // path keeps 128-wide tiles (its fp32 staging tile must fit the pipeline smem)
  int BNsel = bn_select(Nout);
  if (MTsel == 2 && BNsel == 256) BNsel = 128;
  // wave-aware tile width: when 128-wide tiles need more than one wave of 2 x 148 resident CTAs and 256-wide tiles (2 stages,
  // still two CTAs per SM) fit into one with most SMs doubly occupied, the wider tile wins (measured: D's c4 forward
  // 62 -> 52 us, d3's input gradient 36.5 -> 30.4 us); with fewer CTAs the starved 2-stage ring loses (c4 dgrad 58 -> 81 us)
  if (BNsel == 128 && Nout % 256 == 0 && !thin_n && bn256_auto() && MTsel == 1) {
    const long long m_tiles_ = (long long)P.tiles_w * P.tiles_h * tiles_n * g.nclass;
    const long long c128 = m_tiles_ * (Nout / 128), c256 = m_tiles_ * (Nout / 256);
    if (c128 > 296 && c256 <= 296 && c256 >= 200) BNsel = 256;
  }
  int ksplit = 1;
  if (!thin_n && Nout % 64 == 0) {
    const int bn_s = Nout % 128 == 0 ? 128 : 64;
    const long long ctas_s = (long long)P.tiles_w * P.tiles_h * tiles_n * (Nout / bn_s) * g.nclass;
    const long long need_s = (long long)g.N * g.OH * g.OW * Nout * 4;
    P.OH = g.OH; P.OW = g.OW;                 // (partial sums always cover the true output; the finisher crops)
    if (ws && ws_bytes >= need_s && ctas_s <= 74 && g.ntaps * P.kchunks >= 32 && MTsel == 1) {
      while (ksplit * 2 <= g.ntaps && ctas_s * ksplit * 2 <= 296) ksplit *= 2;
      if (ksplit > 1) BNsel = bn_s;
    }
  }
  rc = encode_2d(&P.bmap, wp, K, 16LL * n_rows, thin_n ? 16 : (pair_mode() == 1 && Nout % 256 == 0 && ksplit == 1 && !bn_acc && !ex ? 128 : BNsel));
  if (rc) return rc;

  for (int c = 0; c < g.nclass; ++c)
    for (int j = 0; j < g.ntaps; ++j) {
      const Tap& t = g.tap[c][j];
      if (g.istride == 2) {
        const int p = t.dy & 1, q = t.dx & 1;
        P.tview[c][j] = (int8_t)(p * 2 + q);
        P.tdy[c][j] = (int8_t)((t.dy - p) / 2);
        P.tdx[c][j] = (int8_t)((t.dx - q) / 2);
        // empty parity view (dimension of size 1 in the tensor): such taps only ever read padding
        if ((p == 1 && g.IH < 2) || (q == 1 && g.IW < 2)) { P.tdy[c][j] = 64; P.tdx[c][j] = 64; }
      } else {
        P.tview[c][j] = 0; P.tdy[c][j] = t.dy; P.tdx[c][j] = t.dx;
      }
      P.twt[c][j] = t.wtap;
    }
  (void)geom_kind;
  P.Nout = n_rows;          // row pitch (per tap) of the packed weights
  if (thin_n) {
    dim3 grid((unsigned)(P.tiles_w * P.tiles_h * tiles_n), 1, (unsigned)g.nclass);
    return launch_tapgemm<16, 3>(P, grid, st);    // 54 KB: four CTAs per SM (short K loops: latency-bound)
  }
  const int BN = BNsel;
  // deep-K layers with few output tiles (the U-Net bottleneck): split the taps over extra CTAs, reduce in fp32
  const long long need = (long long)g.N * g.OH * g.OW * Nout * 4;
  if (ksplit > 1) {
    if (!ws_clean) {
      cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)need, st);
      if (e != cudaSuccess) return (int)e;
    }
    P.ksplit = ksplit; P.part_out = ws; P.part_ld = Nout; P.bn_acc = nullptr;
    dim3 grid((unsigned)(P.tiles_w * P.tiles_h * tiles_n), (unsigned)(Nout / BN), (unsigned)(g.nclass * ksplit));
    rc = BN == 256 ? launch_tapgemm<256, 2>(P, grid, st) : BN == 128 ? launch_tapgemm<128, 3>(P, grid, st) : launch_tapgemm<64, 4>(P, grid, st);
    if (rc) return rc;
    const long long Ppix = (long long)g.N * g.OH * g.OW;
    long long blocks = (Ppix * (Nout / 4) + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    if (bn_acc) {
      const int quads = Nout / 4;
      if (quads > 256 || 256 % quads != 0) return STCGAN_EUNSUPPORTED;
      const int rpb = 256 / quads;
      blocks = (Ppix + rpb * 8 - 1) / (rpb * 8); if (blocks > 148) blocks = 148; if (blocks < 1) blocks = 1;
    }
    P.bn_acc = nullptr;     // (the GEMM launch above already ran; statistics come from the finished sums)
    launch_k(splitk_finish_kernel, (unsigned)blocks, 256, 0, st, ws, Ppix, Nout, bias, act, static_cast<__nv_bfloat16*>(y), ldy, bn_acc,
             ws_clean ? 1 : 0, fx);
    return finish_launch();
  }
  if (ex && ex->HC > 0 && ex->WC > 0) { P.OH = ex->HC; P.OW = ex->WC; }
  if (MTsel == 2) {
    if (BNsel == 256) BNsel = 128;
    dim3 grid2((unsigned)(P.tiles_w * P.tiles_h * tiles_n), (unsigned)(Nout / BNsel), (unsigned)g.nclass);
    return BNsel == 128 ? launch_tapgemm<128, 2, 2>(P, grid2, st) : launch_tapgemm<64, 2, 2>(P, grid2, st);
  }
  if (pair_mode() == 1 && Nout % 256 == 0 && ksplit == 1 && !bn_acc && !ex) {   // each CTA stages 128 of the 256 weight rows (TMA box of 128)
    const int m_tiles = P.tiles_w * P.tiles_h * tiles_n;
    return launch_tapgemm_pair<256, 3>(P, m_tiles, Nout / 256, g.nclass, st);
  }
  if (persistent_mode() == 1 && BN != 256 && !bn_acc && !ex) {
    const int m_tiles = P.tiles_w * P.tiles_h * tiles_n;
    if (BN == 128) return launch_tapgemm_persistent<128, 5>(P, m_tiles, Nout / BN, g.nclass, st);
    return launch_tapgemm_persistent<64, 6>(P, m_tiles, Nout / BN, g.nclass, st);
  }
  dim3 grid((unsigned)(P.tiles_w * P.tiles_h * tiles_n), (unsigned)(Nout / BN), (unsigned)g.nclass);
  if (BN == 256) return (bn_select(Nout) == 256 && bn256_stages() == 4) ? launch_tapgemm<256, 4>(P, grid, st) : launch_tapgemm<256, 2>(P, grid, st);
  // launches that leave at most one CTA per SM get a deeper ring instead of a second resident CTA: with 3 stages a lone CTA
  // covers only ~96 KB of the ~170 KB that must be in flight to feed the tensor pipe (measured 0.40 us per 128x128x64 step
  // against 0.21 us with two resident CTAs)
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  const bool deep = ctas <= 148 && g.ntaps * P.kchunks >= 8 && deep_ring_mode();
  // STCGAN_TC_STAGES=n (experiments): force the ring depth (fewer stages = more resident CTAs per SM)
  int force = 0;
  { const char* e = getenv("STCGAN_TC_STAGES"); if (e) force = atoi(e); }
  // STCGAN_TC_DUAL=1 (opt-in): two MMA issuer threads for 128-wide tiles.  Parity-green (tests/test_kernels_gpu.py with the
  // variable set); measured per launch (tools/conv_probe.py): the lone-CTA launches gain (e4 fwd 29.4 -> 24.3 us), launches
  // with two resident CTAs per SM do not (c4 dgrad 57.8 -> 59.9, e3 fwd 19.3 -> 19.6), and ONE of two probe runs ended in an
  // unspecified launch failure that has not been reproduced or explained yet -- hence not a default
  const bool dual = dual_issue_mode() && g.ntaps * P.kchunks >= 4;
  if (BN == 128) {
    if (force == 2) return launch_tapgemm<128, 2>(P, grid, st);
    if (force == 3) return launch_tapgemm<128, 3>(P, grid, st);
    if (dual) return deep ? launch_tapgemm<128, 6, 1, 2>(P, grid, st) : launch_tapgemm<128, 3, 1, 2>(P, grid, st);
    return deep ? launch_tapgemm<128, 6>(P, grid, st) : launch_tapgemm<128, 3>(P, grid, st);
  }
  if (force == 2) return launch_tapgemm<64, 2>(P, grid, st);
  if (force == 3) return launch_tapgemm<64, 3>(P, grid, st);
  if (force == 4) return launch_tapgemm<64, 4>(P, grid, st);
  if (deep) return launch_tapgemm<64, 8>(P, grid, st);
  // 128x64 tiles are cheap (half a 128x128 step per stage) and their launches have thousands of CTAs with short K loops:
  // the per-CTA fixed cost (prologue, first TMA round trip, epilogue) is hidden by residency, not by ring depth --
  // measured on B200 (tools/conv_probe.py): 2048 CTAs x 8 steps 47.7 us with 4 stages / 2 CTAs per SM, 38.5 us with
  // 3 stages / 3 CTAs, 35.3 us with 2 stages / 4 CTAs; 2048 CTAs x 16 steps 68.0 -> 55.4 -> 55.2 us
  if (g.ntaps * P.kchunks <= 8) return launch_tapgemm<64, 2>(P, grid, st);
  return launch_tapgemm<64, 3>(P, grid, st);
}

// 5D im2col view of a zero-bordered 8-channel tensor T [N, HP, WP, 8]:
//   (d0 = (kw, c): 32 contiguous elements = 64 B, d1 = grid column, d2 = grid row, d3 = image, d4 = kernel row kh)
//   element address = ((n*HP + s*gy + kh)*WP + s*gx)*8 + d0
//   box = (32, wt, ht, nt, box_kh) -> box_kh consecutive [pixel][64 B] SWIZZLE_64B blocks from ONE TMA instruction
static int encode_thin5d(CUtensorMap* m, const void* base, long long HP, long long WP, long long N, int s,
                         long long GW, long long GH, int bw, int bh, int bn, int box_kh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t dims[5] = {32, (cuuint64_t)GW, (cuuint64_t)GH, (cuuint64_t)N, 4};
  cuuint64_t strides[4] = {(cuuint64_t)s * 16, (cuuint64_t)s * WP * 16, (cuuint64_t)HP * WP * 16, (cuuint64_t)WP * 16};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)box_kh};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// NHWC bf16 tensor with the channel dimension split in blocks of 64: dims (64, W, H, N, C/64); box (64, bw, bh, bn, nblk)
// -> nblk consecutive [pixel][128 B] SWIZZLE_128B blocks (an MN-major operand of 64*nblk channels) from ONE TMA instruction
static int encode_nhwc_blk(CUtensorMap* m, const void* base, long long C, long long W, long long H, long long N,
                           long long sw, long long sh, long long sn, int bw, int bh, int bn, int nblk) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  if (W < 1) W = 1;
  if (H < 1) H = 1;
  cuuint64_t dims[5] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / 64)};
  cuuint64_t strides[4] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2, 128};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)nblk};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// thin-K convolution: out[p, n] = sum_{kh,kw,c} T[window(p)][kh][kw][c] * Wt[n][(kh*4+kw)*8 + c]
//   T: zero-bordered [N, HP, WP, 8] bf16; output grid OH x OW with window anchor (s*oy, s*ox) in padded coordinates
int thinconv_tc(const void* t, int N, int HP, int WP, int s, const void* wthin, const float* bias, int act,
                void* y, int OH, int OW, int Nout, int ldy, cudaStream_t st, const EpilogueExtra* ex = nullptr) {
  if (Nout % 64 != 0 || ldy % 8 != 0 || !al16(t) || !al16(y) || !al16(wthin)) return STCGAN_EUNSUPPORTED;
  if (act == STCGAN_ACT_TANH || act == STCGAN_ACT_SIGMOID) return STCGAN_EUNSUPPORTED;
  if (ex && ex->y2 && (ex->ldy2 % 8 != 0 || !al16(ex->y2) || ex->act2 == STCGAN_ACT_TANH || ex->act2 == STCGAN_ACT_SIGMOID))
    return STCGAN_EUNSUPPORTED;
  if (s * (OH - 1) + 4 > HP || s * (OW - 1) + 4 > WP) return STCGAN_EINVAL;     // windows must stay inside the border
  TapGemmParams P;
  memset(&P, 0, sizeof(P));
  choose_tile(TC_BM, OW, OH, N, &P.wt, &P.ht, &P.nt);
  P.tiles_w = (OW + P.wt - 1) / P.wt; P.tiles_h = (OH + P.ht - 1) / P.ht;
  const int tiles_n = (N + P.nt - 1) / P.nt;
  P.GH = OH; P.GW = OW; P.N = N; P.OH = OH; P.OW = OW; P.ostride = 1;
  P.ntaps = 1; P.kchunks = 2; P.thin_k = 1;
  P.Nout = Nout; P.nout_real = Nout; P.ldy = ldy; P.act = act; P.bias = bias; P.y = static_cast<__nv_bfloat16*>(y);
  if (ex) { P.ep_scale = ex->scale; P.y2 = static_cast<__nv_bfloat16*>(ex->y2); P.ldy2 = ex->ldy2; P.act2 = ex->act2; }
  int rc = encode_thin5d(&P.amap[0], t, HP, WP, N, s, OW, OH, P.wt, P.ht, P.nt, 2);
  if (rc) return rc;
  for (int v = 1; v < 4; ++v) P.amap[v] = P.amap[0];
  const int BN = Nout % 128 == 0 ? 128 : 64;
  rc = encode_2d(&P.bmap, wthin, 128, Nout, BN);
  if (rc) return rc;
  if (persistent_mode() != 0) {
    const int m_tiles = P.tiles_w * P.tiles_h * tiles_n;
    // two K steps per tile: these launches are bound by the epilogue (TMEM -> bf16 rows -> global), not by the ring, so
    // they run several shallow persistent CTAs per SM (3 x 4 epilogue warps for 128x64 tiles) instead of one deep one
    const char* e = getenv("STCGAN_THINK_DEEP");
    if (e && e[0] == '1') {
      if (BN == 128) return launch_tapgemm_persistent<128, 5>(P, m_tiles, Nout / BN, 1, st);
      return launch_tapgemm_persistent<64, 6>(P, m_tiles, Nout / BN, 1, st);
    }
    if (BN == 128) return launch_tapgemm_persistent<128, 2>(P, m_tiles, Nout / BN, 1, st, 2);
    return launch_tapgemm_persistent<64, 2>(P, m_tiles, Nout / BN, 1, st, 3);
  }
  dim3 grid((unsigned)(P.tiles_w * P.tiles_h * tiles_n), (unsigned)(Nout / BN), 1);
  if (BN == 128) return launch_tapgemm<128, 2>(P, grid, st);
  return launch_tapgemm<64, 2>(P, grid, st);
}

template <int BN, int STAGES>
static int launch_wgrad(const WgradParams& P, dim3 grid, cudaStream_t st) {
  using SM = WgradSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapwgrad_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_k(tapwgrad_tc_kernel<BN, STAGES>, grid, 192, SM::TOTAL, st, P);
  return finish_launch();
}

int tapwgrad_tc(int geom, const void* S, int N, int SH, int SW, int D0, int lds,
                const void* L, int LH, int LW, int D1, int ldl, float* G, cudaStream_t st) {
  if (D0 % 128 != 0 || D1 % 64 != 0 || lds % 8 != 0 || ldl % 8 != 0 || !al16(S) || !al16(L)) return STCGAN_EUNSUPPORTED;
  WgradParams P;
  memset(&P, 0, sizeof(P));
  choose_tile(64, SW, SH, N, &P.wt, &P.ht, &P.nt);
  P.tiles_w = (SW + P.wt - 1) / P.wt; P.tiles_h = (SH + P.ht - 1) / P.ht; P.tiles_n = (N + P.nt - 1) / P.nt;
  P.D0 = D0; P.D1 = D1; P.G = G;
  const int BN = D1 % 128 == 0 ? 128 : 64;
  const int total_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  const int out_tiles = (D0 / 128) * (D1 / BN) * 16;
  // one wave: at most 2 x 148 resident CTAs (a 297th CTA would run alone in a second wave and double the launch time)
  int splits = wgrad_split_floor() ? (2 * 148) / out_tiles : (2 * 148 + out_tiles - 1) / out_tiles;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  P.tiles_per_split = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + P.tiles_per_split - 1) / P.tiles_per_split;

  const __nv_bfloat16* sb = static_cast<const __nv_bfloat16*>(S);
  const __nv_bfloat16* lb = static_cast<const __nv_bfloat16*>(L);
  int rc = encode_nhwc_blk(&P.smap, sb, D0, SW, SH, N, lds, (long long)SW * lds, (long long)SH * SW * lds, P.wt, P.ht, P.nt, 2);
  if (rc) return rc;
  const long long sn = (long long)LH * LW * ldl;
  const int stride = geom == STCGAN_GEOM_WIN_S2 ? 2 : 1;
  if (stride == 1) {
    rc = encode_nhwc_blk(&P.lmap[0], lb, D1, LW, LH, N, ldl, (long long)LW * ldl, sn, P.wt, P.ht, P.nt, BN / 64);
    if (rc) return rc;
    for (int v = 1; v < 4; ++v) P.lmap[v] = P.lmap[0];
  } else {
    for (int p = 0; p < 2; ++p)
      for (int q = 0; q < 2; ++q) {
        const bool empty = (LW - q + 1) / 2 < 1 || (LH - p + 1) / 2 < 1;     // see tapconv_tc: keep the base mapped
        rc = encode_nhwc_blk(&P.lmap[p * 2 + q], empty ? lb : lb + ((long long)p * LW + q) * ldl, D1, (LW - q + 1) / 2,
                             (LH - p + 1) / 2, N, 2LL * ldl, 2LL * LW * ldl, sn, P.wt, P.ht, P.nt, BN / 64);
        if (rc) return rc;
      }
  }
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      const int t = kh * 4 + kw, dy = kh - 1, dx = kw - 1;
      if (stride == 2) {
        const int p = dy & 1, q = dx & 1;
        P.tview[t] = (int8_t)(p * 2 + q); P.tdy[t] = (int8_t)((dy - p) / 2); P.tdx[t] = (int8_t)((dx - q) / 2);
        if ((p == 1 && LH < 2) || (q == 1 && LW < 2)) { P.tdy[t] = 64; P.tdx[t] = 64; }
      } else {
        P.tview[t] = 0; P.tdy[t] = (int8_t)dy; P.tdx[t] = (int8_t)dx;
      }
    }
  dim3 grid((unsigned)((D0 / 128) * (D1 / BN)), 16, (unsigned)splits);
  if (BN == 128) return launch_wgrad<128, 3>(P, grid, st);   // 96 KB: two CTAs per SM (6 stages per SM in flight)
  return launch_wgrad<64, 4>(P, grid, st);
}

// thin weight gradient: D[(tap,c)][d] = sum_q Twin[q,(tap,c)] * F[q,d]  ->  G (see WgradParams)
//   T: zero-bordered [N, HP, WP, 8] bf16 with thin_c real channels; F: [N, FH, FW, Dfat] pitch ldf; the window of T
//   for grid pixel q = (fy, fx) is anchored at (s*fy, s*fx) in padded coordinates.
int thinwgrad_tc(const void* t, int N, int HP, int WP, int s, int thin_c, const void* f, int FH, int FW, int Dfat, int ldf,
                 int fat_is_dim0, int flip, float* G, cudaStream_t st) {
  if (Dfat % 64 != 0 || ldf % 8 != 0 || !al16(t) || !al16(f) || thin_c < 1 || thin_c > 8) return STCGAN_EUNSUPPORTED;
  if (s * (FH - 1) + 4 > HP || s * (FW - 1) + 4 > WP) return STCGAN_EINVAL;
  WgradParams P;
  memset(&P, 0, sizeof(P));
  choose_tile(64, FW, FH, N, &P.wt, &P.ht, &P.nt);
  P.tiles_w = (FW + P.wt - 1) / P.wt; P.tiles_h = (FH + P.ht - 1) / P.ht; P.tiles_n = (N + P.nt - 1) / P.nt;
  P.thin = 1; P.thin_c = thin_c; P.fat_is_dim0 = fat_is_dim0; P.flip = flip; P.Dfat = Dfat;
  P.D0 = 128; P.D1 = Dfat; P.G = G;
  const int BN = Dfat % 128 == 0 ? 128 : 64;
  const int total_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  const int out_tiles = Dfat / BN;
  int splits = wgrad_split_floor() ? (2 * 148) / out_tiles : (2 * 148 + out_tiles - 1) / out_tiles;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  P.tiles_per_split = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + P.tiles_per_split - 1) / P.tiles_per_split;
  int rc = encode_nhwc_blk(&P.smap, f, Dfat, FW, FH, N, ldf, (long long)FW * ldf, (long long)FH * FW * ldf, P.wt, P.ht, P.nt, BN / 64);
  if (rc) return rc;
  rc = encode_thin5d(&P.lmap[0], t, HP, WP, N, s, FW, FH, P.wt, P.ht, P.nt, 4);
  if (rc) return rc;
  dim3 grid((unsigned)out_tiles, 1, (unsigned)splits);
  if (BN == 128) return launch_wgrad<128, 3>(P, grid, st);   // 96 KB: two CTAs per SM (6 stages per SM in flight)
  return launch_wgrad<64, 4>(P, grid, st);
}

}  // namespace stcgan

/*
 * stcgan_b200 -- C ABI of the B200-native ST-CGAN hot path.
 *
 * The reference (nhchiu/Shadow-Removal-ISTD) is pure Python and has NO FFI/plugin layer:
 * its hot path calls torch.nn ops which dispatch to cuDNN/oneDNN.
 * The drop-in seam is therefore the nn.Module contract (python package `stcgan_b200`);
 * this header is the boundary *below* that seam: every piece of arithmetic the reference
 * obtains from torch on this path is an entry point here, on raw device pointers.
 * Each entry point cites the reference call site(s) whose arithmetic it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `host`;
 *   - activations are NHWC ("pixel-major"): element (n,h,w,c) of a tensor with pixel pitch
 *     `ld` lives at ((n*H + h)*W + w)*ld + c; a channel slice of a wider buffer is addressed by
 *     offsetting the base pointer and keeping the wide `ld` (that is how U-Net skip
 *     concatenations are produced in place instead of by a copy);
 *   - `dtype` selects the activation/packed-weight element type: STCGAN_F32 or STCGAN_BF16;
 *     accumulation is always fp32 (statistics: fp64);
 *   - packed conv weights are tap-major: Wp[t][n][k], t = kh*4+kw, n = output channel of the
 *     GEMM, k = reduction channel;  packed weight gradients are G[t][d0][d1] fp32 where
 *     (d0,d1) are the first two dims of the torch parameter ([Cout,Cin,4,4] for Conv2d,
 *     [Cin,Cout,4,4] for ConvTranspose2d);
 *   - every function returns 0 on success, a negative STCGAN_E* code on bad arguments, or a
 *     positive cudaError_t; nothing throws or exits across this boundary;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - nothing here allocates or retains caller memory; the caller (torch's caching allocator
 *     in the Python host) owns every buffer.
 */
#ifndef STCGAN_B200_H
#define STCGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STCGAN_ABI_VERSION 1

enum { STCGAN_F32 = 0, STCGAN_BF16 = 1 };

/* error codes */
enum {
  STCGAN_OK = 0,
  STCGAN_EINVAL = -1,      /* bad argument (shape / alignment / enum) */
  STCGAN_EUNSUPPORTED = -2 /* shape not supported by the selected backend */
};

/* activation codes (fused into epilogues / BN apply) */
enum { STCGAN_ACT_NONE = 0, STCGAN_ACT_LEAKY = 1, STCGAN_ACT_RELU = 2, STCGAN_ACT_TANH = 3, STCGAN_ACT_SIGMOID = 4 };

/* gather geometries of the implicit GEMM  out[p, n] = sum_t sum_k in[gather_t(p), k] * Wp[t][n][k] */
enum {
  STCGAN_GEOM_WIN_S2 = 0,      /* 4x4 window, stride 2, pad 1 : Conv2d(4,2,1) forward; ConvTranspose2d(4,2,1) dgrad */
  STCGAN_GEOM_WIN_S1 = 1,      /* 4x4 window, stride 1, pad 1 : Conv2d(4,1,1) forward */
  STCGAN_GEOM_WIN_S1_FLIP = 2, /* 4x4 window, stride 1, pad 2, flipped taps : Conv2d(4,1,1) dgrad */
  STCGAN_GEOM_PARITY = 3       /* 4 output-parity classes x 2x2 taps : ConvTranspose2d(4,2,1) forward; Conv2d(4,2,1) dgrad */
};

/* compute back-ends */
enum {
  STCGAN_BACKEND_FFMA = 0, /* CUDA-core fp32-accumulate tiles: any shape, both dtypes (the fp32 mode and thin layers) */
  STCGAN_BACKEND_TC = 1    /* tcgen05 + TMEM + TMA implicit GEMM, bf16 only, K%64==0, Nout%64==0 */
};

/* ---- library ---------------------------------------------------------------------------- */
int stcgan_abi_version(void);
/* compiled arch string, e.g. "sm_100a" */
const char* stcgan_arch(void);
/* human-readable text for a return code of any function below */
const char* stcgan_error_string(int code);
/* number of kernels launched by this library since the last reset (host counter) */
int64_t stcgan_launch_count(void);
void stcgan_launch_count_reset(void);

/* ---- convolutions as tap-GEMMs ------------------------------------------------------------
 * replaces: nn.Conv2d(k=4,s=2,p=1) forward/backward  src/models/stcgan_g.py:85-86, stcgan_d.py:22-23,33-35
 *           nn.Conv2d(k=4,s=1,p=1) forward/backward  src/models/stcgan_d.py:43-44,49-50
 *           nn.ConvTranspose2d(k=4,s=2,p=1) fwd/bwd   src/models/stcgan_g.py:93-95,100-102,107-109
 * (the reference reaches cuDNN/oneDNN through ATen for all of them).
 *
 * x    : input  [N, IH, IW, K]   pitch ldx      (dtype)
 * wp   : packed weights [16][Nout][K]            (dtype)
 * bias : optional fp32 [Nout] (NULL = none)
 * y    : output [N, OH, OW, Nout] pitch ldy      (dtype), or, if out_nchw_f32 != 0, a float tensor
 *        [N, Nout, OH, OW] (used for the generator's final Tanh output)
 * act  : epilogue activation applied after bias
 * IH/IW may be smaller than the window reach implies: out-of-range taps read zeros, which is
 * also how the reference's odd-size F.pad (stcgan_g.py:126-132) is realised without a copy.
 * workspace (optional, fp32, >= N*OH*OW*Nout*4 bytes): lets the tensor-core backend split the taps of deep-K layers with
 * few output tiles (the U-Net bottleneck) over more CTAs; partial sums are reduced there in fp32.
 * workspace_bytes < 0: the workspace (|workspace_bytes| bytes) is all zeros on entry and is left all zeros on exit (the
 * finishing kernel clears what it reads), which saves the memset launch in front of every such convolution.
 */
int stcgan_tapconv(int geom, int dtype, int backend,
                   const void* x, int N, int IH, int IW, int K, int ldx,
                   const void* wp, const float* bias, int act,
                   void* y, int OH, int OW, int Nout, int ldy, int out_nchw_f32,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* The same convolution with the following BatchNorm's batch statistics fused into the GEMM epilogue (tensor-core backend,
 * bf16, act == STCGAN_ACT_NONE, Nout % 64 == 0): per output channel the sum and the sum of squares of the bf16-rounded
 * outputs are added (fp64 atomics) to bn_acc[STCGAN_BN_SLOTS][2][Nout] -- STCGAN_BN_SLOTS partial slots that spread the
 * atomic traffic; the caller zeroes bn_acc and stcgan_bn_fused_apply sums the slots.  Replaces the separate statistics pass
 * over the conv output of nn.BatchNorm2d in training mode (stcgan_g.py:88,90; stcgan_d.py:36,45). */
#define STCGAN_BN_SLOTS 4
int stcgan_tapconv_bnstats(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp,
                           void* y, int OH, int OW, int Nout, int ldy, void* workspace, int64_t workspace_bytes,
                           double* bn_acc, void* stream);
/* Inference form of the tensor-core convolution (bf16): eval-mode BatchNorm folded into the epilogue and up to two
 * activated outputs -- replaces nn.Conv2d / nn.ConvTranspose2d + nn.BatchNorm2d(eval) + nn.LeakyReLU / nn.ReLU
 * (src/models/stcgan_g.py:85-90,107-111 as executed by CGAN.infer, src/cgan.py:422-438) without materialising the
 * pre-activation tensor:  v = acc * scale[c] + shift[c]  (scale NULL = 1, shift NULL = 0; for BatchNorm scale =
 * gamma/sqrt(var+eps), shift = beta - mean*scale, i.e. stcgan_bn_finalize's scale_shift);  y = act(v);  y2 = act2(v) (y2 NULL
 * = none).  Stores are cropped to [HC, WC] <= [OH, OW] (the odd-size crop of stcgan_g.py:131; y / y2 are [N, HC, WC, *]
 * views with pitches ldy / ldy2).  act, act2 in {NONE, LEAKY, RELU}.  Workspace as stcgan_tapconv. */
int stcgan_tapconv_ep(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp,
                      const float* scale, const float* shift, int act, void* y, int ldy, int act2, void* y2, int ldy2,
                      int OH, int OW, int HC, int WC, int Nout, void* workspace, int64_t workspace_bytes, void* stream);

/* weight gradient of the same convolutions:  G[t][d0][d1] += sum_p S[p, d0] * L[win_t(p), d1]
 * S : "small-grid" tensor [N, SH, SW, D0] pitch lds (Conv2d: dY; ConvTranspose2d: the layer input)
 * L : "large-grid" tensor [N, LH, LW, D1] pitch ldl (Conv2d: the layer input; ConvTranspose2d: dY)
 * geom : STCGAN_GEOM_WIN_S2 or STCGAN_GEOM_WIN_S1 (window of the *forward* Conv2d, or of the
 *        ConvTranspose2d seen from its output side)
 * G accumulates (callers zero it once per backward phase), fp32. */
int stcgan_tapwgrad(int geom, int dtype, int backend,
                    const void* S, int N, int SH, int SW, int D0, int lds,
                    const void* L, int LH, int LW, int D1, int ldl,
                    float* G, void* stream);

/* ---- thin layers on the tensor cores (bf16) --------------------------------------------------------
 * The first / last layers (Cin in {3,4,7}: stcgan_g.py:85-86 outermost, stcgan_d.py:22-23; Cout in {1,3}:
 * stcgan_g.py:93-95, stcgan_d.py:49-50) have one GEMM dimension below a tensor-core tile.  Three layouts keep them on
 * tcgen05 without inflating HBM traffic:
 *   thin N : (a) stcgan_thin_col2im below -- the form the networks use: one pixel GEMM with the taps in N + in-CTA col2im;
 *            (b) stcgan_tapconv_thin_n -- the earlier tap-GEMM form, kept for comparison (the networks no longer call it):
 *            N padded to 16 in the packed weights only (stcgan_pack_weight_pad16); output written straight to NCHW fp32
 *            (bias + Tanh/Sigmoid fused) or to an 8-channel NHWC gradient tensor;
 *   thin K : the thin tensor is kept zero-bordered with 8 channels, so one 4x4xC window is 4 rows of 64 contiguous
 *            bytes; a 5-D TMA view turns it into two 128-byte K-chunks per output pixel (K = 128), weights packed
 *            [n][(kh*4+kw)*8+c] (stcgan_pack_weight_thin);
 *   thin wgrad : the same 5-D view as the M operand (128 = 16 taps x 8 channels), the fat tensor as N.
 */
int stcgan_tapconv_thin_n(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp16,
                          const float* bias, int act, void* y_nhwc8, int ldy, float* y_nchw_f32, int OH, int OW, int Nout,
                          void* stream);
/* out[n,oy,ox,:] = act(bias + sum_{kh,kw,c} t[n, stride*oy+kh, stride*ox+kw, c] * wthin[:, (kh*4+kw)*8+c]);
 * t is zero-bordered [N, HP, WP, 8] bf16 (the border carries the conv padding), wthin [Nout][128] bf16 */
int stcgan_thinconv(const void* t, int N, int HP, int WP, int stride, const void* wthin, const float* bias, int act,
                    void* y, int OH, int OW, int Nout, int ldy, void* stream);
/* the same convolution with TWO activated outputs of the one accumulator: y = act(v), y2 = act2(v).  This is the U-Net's
 * first layer: its raw output feeds LeakyReLU (next down conv) and, through the skip concatenation, ReLU (the decoder) --
 * src/models/stcgan_g.py:87,89,125 -- and neither needs the pre-activation tensor (sign(leaky(v)) == sign(v) serves the
 * backward pass), so it is never written. */
int stcgan_thinconv2(const void* t, int N, int HP, int WP, int stride, const void* wthin, const float* bias, int act,
                     void* y, int ldy, int act2, void* y2, int ldy2, int OH, int OW, int Nout, void* stream);
/* G += sum_q window(t, q)[(tap,c)] * f[q, d]; G index = [wtap][d][c] if fat_is_dim0 else [wtap][c][d] with
 * wtap = flip ? 15 - tap : tap; t zero-bordered [N, HP, WP, 8] with thin_c real channels, f [N, FH, FW, Dfat] pitch ldf */
int stcgan_thinwgrad(const void* t, int N, int HP, int WP, int stride, int thin_c, const void* f, int FH, int FW, int Dfat,
                     int ldf, int fat_is_dim0, int flip, float* G, void* stream);
/* Thin-N convolutions as one pixel GEMM + in-CTA col2im (thin_col2im.cu): the fat input x [N, IH, IW, K] (pitch ldx) is read
 * ONCE; Pm[pixel][(tap,c)] = x[pixel,:] . wt[(tap,c),:] on tcgen05 (N dimension = 16 taps x cpad channels), and the 16 tap
 * planes are summed inside the CTA over overlapping pixel tiles, so every output is produced by exactly one CTA (no atomics).
 *   mode 0: stride-2 scatter = nn.ConvTranspose2d(k4,s2,p1) forward (stcgan_g.py:93-95) and the input gradient of
 *           nn.Conv2d(k4,s2,p1) (first layers, stcgan_g.py:85-86 outermost, stcgan_d.py:22-23);  OH <= 2*IH+1, OW <= 2*IW+1
 *   mode 1: stride-1 gather = nn.Conv2d(k4,s1,p1) forward (stcgan_d.py:49-50);  OH = IH - 1, OW = IW - 1
 * wt: [16*cpad][K] bf16 from stcgan_pack_weight_tapn, cpad in {1,4,8}, cout <= cpad real channels.
 * Output: y_nchw_f32 [N, cout, OH, OW] with bias + activation (any STCGAN_ACT_*), or (mode 0 only, no bias / activation)
 * y_nhwc8: bf16 [N, OH, OW, >= 8] pitch ldy, channels cout..7 written as zeros.  Exactly one of the two is non-NULL. */
int stcgan_thin_col2im(int mode, const void* x, int N, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                       const float* bias, int act, float* y_nchw_f32, void* y_nhwc8, int ldy, int OH, int OW, void* stream);
/* The generators' last layer for CGAN.infer (src/cgan.py:437-446): ConvTranspose2d(128 -> cout <= 8, k4 s2 p1) + bias + Tanh as
 * above (mode 0), with the image quantisation of utils.float2uint (src/utils.py:65-67) INSIDE the same epilogue:
 * y_nhwc_u8[n,oy,ox,c] = (uint8) trunc(clip(act(v) * 0.5 + 0.5, 0, 1) * 255), float32 arithmetic like numpy (bit-exact with
 * quantising the float output afterwards).  y_nchw_f32 may be NULL when only the image is wanted. */
int stcgan_thin_convt_u8(const void* x, int N, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                         const float* bias, int act, float* y_nchw_f32, uint8_t* y_nhwc_u8, int OH, int OW, void* stream);
/* wt[(t*cpad + r)][k] = W(r, k, t) with (r, k) = (d0, d1) if n_is_d0 else (d1, d0), zero rows for r >= the thin dimension */
int stcgan_pack_weight_tapn(const float* w, int D0, int D1, int n_is_d0, int cpad, void* out, void* stream);
/* thin weight packings from the torch parameter W[d0][d1][4][4] (fp32) to bf16 */
int stcgan_pack_weight_thin(const float* w, int D0, int D1, int n_is_d0, int flip, void* out, void* stream);
int stcgan_pack_weight_pad16(const float* w, int D0, int D1, int n_is_d0, void* out, void* stream);

/* ---- weight layout ---------------------------------------------------------------------------
 * torch parameter W[d0][d1][4][4] fp32  ->  P1[t][d0][d1] and P2[t][d1][d0] in `dtype`
 * (either output may be NULL).  State-dict layout: src/models/stcgan_g.py:85-109, stcgan_d.py:22-50. */
int stcgan_pack_weight(int dtype, const float* w, int D0, int D1, void* p1, void* p2, void* stream);
/* packed gradient G[t][d0][d1] fp32 -> torch layout grad[d0][d1][16]; accumulate != 0 adds */
int stcgan_unpack_grad(const float* g, int D0, int D1, float* grad, int accumulate, void* stream);

/* ---- BatchNorm2d + activation -----------------------------------------------------------------
 * replaces: nn.BatchNorm2d (train/eval) + nn.LeakyReLU(0.2,True) / nn.ReLU(True)
 *           src/models/stcgan_g.py:87-90, stcgan_d.py:24,36-37,45-46
 */
/* per-channel sum / sum-of-squares of y [P pixels x C] (pitch ld) accumulated into fp64 acc[2][C] (caller zeroes) */
int stcgan_bn_stats(int dtype, const void* y, int64_t P, int C, int ld, double* acc, void* stream);
/* training: mean/biased var from acc over count P -> mean_invstd[2][C], scale_shift[2][C] (scale = gamma*invstd,
 * shift = beta - mean*scale); running_mean/var momentum update with unbiased var (NULL = skip).
 * eval (training == 0): statistics taken from running_mean/var, acc ignored. */
int stcgan_bn_finalize(const double* acc, int64_t P, int C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int training,
                       float* mean_invstd, float* scale_shift, void* stream);
/* the running-statistics half of the training-mode call above, on its own (nn.BatchNorm2d's momentum update,
 * src/models/stcgan_d.py:36,45): statistics = sum over the STCGAN_BN_SLOTS slots of acc[slot][2][C] over `count` values per
 * channel; running_mean / running_var <- (1 - momentum) * old + momentum * (mean / unbiased variance).  Used after a forward
 * pass that ran stcgan_bn_fused_apply with running_mean == NULL, so that two forward passes of one network (cgan.py:321-324:
 * D(real) and D(fake)) can execute concurrently while the updates still land in the reference's order. */
int stcgan_bn_running_update(const double* acc, int64_t count, int C, float* running_mean, float* running_var,
                             float momentum, void* stream);
/* out1 = act1(y*scale + shift) (and optionally out2 = act2(..)) over the cropped region [N, HC, WC] of
 * y [N, H, W, C]; scale_shift == NULL means identity (layers without BatchNorm).  out2 may be NULL. */
int stcgan_bn_act_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy,
                        const float* scale_shift, int HC, int WC,
                        void* out1, int ld1, int act1, void* out2, int ld2, int act2, void* stream);
/* stcgan_bn_finalize + stcgan_bn_act_apply in ONE launch.  training != 0: statistics = sum over the STCGAN_BN_SLOTS slots
 * of acc[slot][2][C] (stcgan_bn_stats fills slot 0, stcgan_tapconv_bnstats all of them) over `count` values per channel;
 * training == 0: running statistics.  Writes mean_invstd / scale_shift (for the backward pass), updates the running
 * statistics (training, NULL = skip) and applies out1 = act1(y*scale+shift) [, out2 = act2(..)] over the crop. */
int stcgan_bn_fused_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy,
                          const double* acc, int64_t count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, int training,
                          float* mean_invstd, float* scale_shift, int HC, int WC,
                          void* out1, int ld1, int act1, void* out2, int ld2, int act2, void* stream);
/* training-mode BatchNorm(+activation) backward of a SMALL tensor (N*H*W*C <= 64*512 elements, bf16: the innermost U-Net
 * levels) as ONE single-block launch: reduce, coefficients and apply of the two functions below in one kernel, same
 * arithmetic.  Returns STCGAN_EUNSUPPORTED for larger tensors / other dtypes (run the two-pass form then).  dgamma / dbeta
 * are ACCUMULATED into (NULL = skip). */
int stcgan_bn_act_bwd_small(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* scale_shift,
                            const float* mean_invstd, const float* gamma, int HC, int WC, const void* g1, int ldg1, int act1,
                            const void* g2, int ldg2, int act2, void* dy, int lddy, float* dgamma, float* dbeta, void* stream);
/* backward, pass 1: dz = g1*act1'(z) + g2*act2'(z) (z = y*scale+shift, zero outside the crop);
 * acc[slot][0][c] += sum dz, acc[slot][1][c] += sum dz*(y-mean)   (fp64; acc is [STCGAN_BN_SLOTS][2][C], caller zeroes;
 * the blocks spread their atomics over the slots) */
int stcgan_bn_act_bwd_reduce(int dtype, const void* y, int N, int H, int W, int C, int ldy,
                             const float* scale_shift, const float* mean_invstd, int HC, int WC,
                             const void* g1, int ldg1, int act1, const void* g2, int ldg2, int act2,
                             double* acc, void* stream);
/* backward, pass 2: dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat))  (training; the means come from the sum of
 *                   the slots of acc);  dy = gamma*invstd*dz (eval) ;  dy = dz (scale_shift == NULL: no BatchNorm)
 * also dgamma += invstd*sum(acc[.][1]), dbeta += sum(acc[.][0]) (by the first block; NULL = skip), and, if dbias != NULL,
 * dbias[c] += sum_p dy[p,c] (the bias gradient of a preceding biased conv: stcgan_d.py:22-23). */
int stcgan_bn_act_bwd_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy,
                            const float* scale_shift, const float* mean_invstd, const float* gamma,
                            int training, int HC, int WC,
                            const void* g1, int ldg1, int act1, const void* g2, int ldg2, int act2,
                            const double* acc, void* dy, int lddy, float* dgamma, float* dbeta, float* dbias,
                            void* stream);
/* column sums of g [P x C] (pitch ld) added to fp32 out[C]  (bias gradients, stcgan_g.py:93-95, stcgan_d.py:22-23,49-50) */
int stcgan_colsum(int dtype, const void* g, int64_t P, int C, int ld, float* out, void* stream);

/* ---- tensor layout at the module boundary ------------------------------------------------------
 * replaces the torch.cat of NCHW inputs (src/cgan.py:281-289, 321-324): up to three NCHW fp32 sources are
 * gathered into one NHWC tensor of Cpad >= c0+c1+c2 channels (extra channels zero). */
int stcgan_pack_input(int dtype, const float* s0, int c0, const float* s1, int c1, const float* s2, int c2,
                      int N, int H, int W, int border, void* out, int Cpad, void* stream);
/* `border` > 0 writes a zero frame of that many pixels around every image: out is [N, H+2b, W+2b, Cpad].  The thin
 * tensor-core kernels below read their 4x4 windows out of such zero-bordered 8-channel tensors. */
/* gradient of the above for a channel range: grad_nchw[n, c, h, w] (+)= g[n,h,w, coff + c] , c < cn */
int stcgan_unpack_input_grad(int dtype, const void* g, int N, int H, int W, int ldg, int coff, int cn,
                             float* grad_nchw, int accumulate, void* stream);
/* NHWC (dtype) -> NCHW fp32 (module outputs, e.g. discriminator logits) and back */
int stcgan_nhwc_to_nchw(int dtype, const void* x, int N, int H, int W, int C, int ld, float* out, void* stream);
int stcgan_nchw_to_nhwc(int dtype, const float* x, int N, int H, int W, int C, void* out, int ld, void* stream);
/* g_nhwc = dout * (1 - out^2) (Tanh backward, stcgan_g.py:97) or dout*out*(1-out) (Sigmoid, stcgan_d.py:52-53), NCHW fp32 in */
int stcgan_out_act_bwd(int dtype, int act, const float* out_nchw, const float* dout_nchw, int N, int H, int W, int C,
                       int border, void* g, int ldg, void* stream);
/* with border > 0, g is [N, H+2b, W+2b, ldg] zero-bordered and zero in channels >= C */

/* ---- losses ---------------------------------------------------------------------------------------
 * replaces: AdversarialLoss.cal_loss (src/loss.py:79-84: ls==0 -> MSE, ls!=0 -> BCE-with-logits against a scalar
 * target) and DataLoss (src/loss.py:25-26: L1 mean), forward value AND gradient in one pass, for up to 8 terms
 * in one launch.  term kinds: 0 = L1(a, b), 1 = MSE(a, target), 2 = BCE-with-logits(a, target).
 * loss_out[slot] += loss_weight * mean(term);  grad (optional) (+)= weight * d mean(term) / d a.
 */
typedef struct stcgan_loss_term {
  const float* a;      /* prediction / logits, fp32, n elements (any layout: elementwise) */
  const float* b;      /* L1 target (kind 0), else ignored */
  float* grad;         /* d/da, may be NULL */
  int64_t n;
  float target;        /* scalar label for kinds 1,2 */
  float weight;        /* gradient weight: grad (+)= weight * d mean(term)/da   (lambda * 0.5 etc.) */
  int32_t kind;
  int32_t slot;        /* index into loss_out */
  int32_t accumulate;  /* grad += instead of = */
  float loss_weight;   /* value weight: loss_out[slot] += loss_weight * mean(term) */
} stcgan_loss_term;
int stcgan_fused_loss(const stcgan_loss_term* host_terms, int nterms, float* loss_out, void* stream);

/* relativistic logits of AdversarialLoss (src/loss.py:88-96, 102-110) on [N, M] float tensors (M = elements per sample):
 *   backward == 0: out[n,i] = a[n,i] - b[n,i]  (avg == 0, RpGAN)  or  a[n,i] - mean over the batch of b[.,i]  (avg != 0, RaGAN)
 *   backward != 0: the gradient of that map w.r.t. b for an upstream gradient passed as `b`: out = -b  or  -mean over the batch
 *                  of b[.,i] broadcast to every n (a is ignored) */
int stcgan_rel_logits(const float* a, const float* b, int N, int64_t M, int avg, int backward, float* out, void* stream);

/* ---- optimiser --------------------------------------------------------------------------------------
 * replaces torch.optim.Adam(betas, eps=1e-8, no weight decay, no amsgrad) of src/cgan.py:85-90,305,351 as one
 * multi-tensor launch.  Gradients may be in packed [16][d0][d1] layout (d0 > 0) or torch layout (d0 == 0).
 * grad_scale multiplies g (1/world for DDP). */
typedef struct stcgan_adam_tensor {
  float* p; const float* g; float* m; float* v;
  int64_t n;
  int32_t d0, d1;     /* packed-gradient dims, 0 = gradient in parameter layout */
  void* p1; void* p2; /* optional (both or neither; needs d0 % T == 0 and d1 % T == 0 with T = stcgan_adam_tile(), and
                         16-byte aligned p, g, m, v, p1, p2): bf16 tap-major copies P1[t][d0][d1], P2[t][d1][d0] refreshed
                         from the updated parameters in the same pass.  Such a tensor is covered by (d0/T)*(d1/T) blocks
                         (chunk = tile index, d1-tiles fastest) instead of ceil(n / stcgan_adam_chunk()) */
} stcgan_adam_tensor;
/* the table lives in DEVICE memory (built once); blocks[] maps each CUDA block to (tensor, chunk).
 * dev_hyper is a DEVICE array of 8 floats, in/out: {lr, beta1, beta2, eps, grad_scale, steps_done, -, -}.
 * The call first increments steps_done ON THE DEVICE and derives the bias corrections from it (slots 6,7), then
 * updates all tensors -- so a captured CUDA graph of the whole train step replays correctly without any
 * host-side refresh; the host rewrites lr / steps_done only when the schedule or a checkpoint changes them. */
int stcgan_adam_step(const stcgan_adam_tensor* dev_table, const int32_t* dev_blocks, int nblocks,
                     float* dev_hyper, void* stream);
/* the same update for the blocks [first_block, first_block + nblocks) of dev_blocks only (e.g. the tensors of one network, so
 * that its update can start while another network's backward pass is still running).  tick != 0 advances steps_done and the
 * bias corrections: set it on exactly one partial launch per optimiser step, ordered before the others.
 * max_ctas > 0 caps the grid (the CTAs then loop over the blocks): a launch that runs underneath compute-bound kernels of
 * another stream should not take every SM's shared memory. */
int stcgan_adam_step_range(const stcgan_adam_tensor* dev_table, const int32_t* dev_blocks, int first_block, int nblocks,
                           float* dev_hyper, int tick, int max_ctas, void* stream);
/* elements per block chunk used by stcgan_adam_step (host helper for building dev_blocks) */
int stcgan_adam_chunk(void);
/* tile edge T (in (d0, d1) pairs) of tensors whose packed bf16 copies are refreshed by stcgan_adam_step */
int stcgan_adam_tile(void);

/* ---- inference post-processing ------------------------------------------------------------------------
 * replaces `*0.5+0.5` (src/cgan.py:441-442) + utils.float2uint (src/utils.py:65-67) + CHW->HWC transpose
 * (cgan.py:443-446): u8[n,h,w,c] = (uint8) trunc(clip(v*0.5+0.5, 0, 1) * 255), all in fp32 like numpy. */
int stcgan_float2uint_hwc(const float* nchw, int N, int C, int H, int W, uint8_t* out_nhwc, void* stream);
/* ---- input pre-processing on the GPU ----------------------------------------------------------------
 * replaces utils.uint2float (src/utils.py:60-62: astype(float32)/255) + the ISTDDataset normalisation and HWC->CHW
 * transpose (src/dataset.py:152: (s.transpose(2,0,1) - 0.5) * 2): uint8 [N,H,W,C] -> float32 [N,C,H,W], bit-exact with
 * numpy's float32 arithmetic, so the host only ships the decoded uint8 images (4x fewer H2D bytes). */
int stcgan_u8_hwc_to_nchw_f32(const uint8_t* in_nhwc, int N, int H, int W, int C, float* out_nchw, void* stream);
/* The reference's training augmentation on the GPU, fused with the transform above (src/transform.py:57-156 as composed by
 * src/cgan.py:105-110: RandomScale -> RandomRotate -> RandomHorizontalFlip -> RandomCrop, applied to utils.uint2float(image)):
 * uint8 [N,H,W,C] (C = 1 or 3) -> float32 [N,C,crop_h,crop_w] in [-1,1].  The host draws the random numbers (in the
 * reference's order: stcgan_b200.augment.sample_params) and passes, per image, the INVERSES of the two cv.warpAffine
 * matrices in float64 (formed like cv::warpAffine forms them), the flip flag and the crop offsets; both warps are bilinear
 * with a zero border, the second one resampling the first one's output as in the reference.  crop <= image size only. */
typedef struct stcgan_aug_sample {
  double scale_inv[6];   /* row-major 2x3: source position = scale_inv * (x, y, 1) */
  double rot_inv[6];
  int32_t flip, row_off, col_off, identity;   /* identity != 0: scale == 1 and angle == 0 */
} stcgan_aug_sample;
int stcgan_augment_u8(const uint8_t* img_nhwc, int N, int H, int W, int C, const stcgan_aug_sample* dev_samples,
                      int crop_h, int crop_w, float* out_nchw, void* stream);
/* plain float2uint on a flat array (known-answer tests) */
int stcgan_float2uint(const float* in, int64_t n, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STCGAN_B200_H */

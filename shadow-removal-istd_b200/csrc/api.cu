#include <cstdlib>
// extern "C" surface of libstcgan_b200.so (see include/stcgan_b200.h for the contract).
#include "common.cuh"

namespace stcgan {
std::atomic<int64_t> g_launches{0};

int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("STCGAN_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

// tapconv_ffma.cu
int tapconv_ffma(const Geom& g, int dtype, const void* x, int K, int ldx, const void* wp, const float* bias, int act,
                 void* y, int Nout, int ldy, int out_nchw_f32, cudaStream_t st);
int tapwgrad_ffma(int geom, int dtype, const void* S, int N, int SH, int SW, int D0, int lds,
                  const void* L, int LH, int LW, int D1, int ldl, float* G, cudaStream_t st);
// tapconv_tc.cu
int tapconv_tc(int geom_kind, const Geom& g, const void* x, int K, int ldx, const void* wp, const float* bias, int act,
               void* y, int Nout, int ldy, cudaStream_t st, int thin_n, float* y32, float* ws, long long ws_bytes,
               double* bn_acc, const EpilogueExtra* ex);
int thinconv_tc(const void* t, int N, int HP, int WP, int s, const void* wthin, const float* bias, int act,
                void* y, int OH, int OW, int Nout, int ldy, cudaStream_t st, const EpilogueExtra* ex);
int thinwgrad_tc(const void* t, int N, int HP, int WP, int s, int thin_c, const void* f, int FH, int FW, int Dfat, int ldf,
                 int fat_is_dim0, int flip, float* G, cudaStream_t st);
int tapwgrad_tc(int geom, const void* S, int N, int SH, int SW, int D0, int lds,
                const void* L, int LH, int LW, int D1, int ldl, float* G, cudaStream_t st);
// thin_col2im.cu
int thin_col2im_tc(int mode, const void* x, int NB, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                   const float* bias, int act, float* y32, void* y8, int ldy, int OH, int OW, cudaStream_t st, uint8_t* u8);
int pack_weight_tapn(const float* w, int D0, int D1, int n_is_d0, int cpad, void* out, cudaStream_t st);
// bn_act.cu
int bn_stats(int dtype, const void* y, long long P, int C, int ld, double* acc, cudaStream_t st);
int bn_finalize(const double* acc, long long P, int C, const float* gamma, const float* beta, float* rmean, float* rvar,
                float momentum, float eps, int training, float* mean_invstd, float* scale_shift, cudaStream_t st);
int augment_u8(const uint8_t* img, int N, int H, int W, int C, const void* samples, int CH, int CW, float* out, cudaStream_t st);
int bn_running_update(const double* acc, long long count, int C, float* rmean, float* rvar, float momentum, cudaStream_t st);
int bn_act_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, int HC, int WC,
                 void* o1, int ld1, int act1, void* o2, int ld2, int act2, cudaStream_t st);
int bn_act_bwd_reduce(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                      int HC, int WC, const void* g1, int ldg1, int act1, const void* g2, int ldg2, int act2,
                      double* acc, cudaStream_t st);
int bn_act_bwd_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                     const float* gamma, int training, int HC, int WC, const void* g1, int ldg1, int act1,
                     const void* g2, int ldg2, int act2, const double* acc, void* dy, int lddy,
                     float* dgamma, float* dbeta, float* dbias, cudaStream_t st);
int colsum(int dtype, const void* g, long long P, int C, int ld, float* out, cudaStream_t st);
int bn_act_bwd_small(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* ss, const float* mi,
                     const float* gamma, int HC, int WC, const void* g1, int ldg1, int act1, const void* g2, int ldg2,
                     int act2, void* dy, int lddy, float* dgamma, float* dbeta, cudaStream_t st);
int bn_fused_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const double* acc, long long count,
                   const float* gamma, const float* beta, float* rmean, float* rvar, float momentum, float eps, int training,
                   float* mean_invstd, float* scale_shift, int HC, int WC,
                   void* o1, int ld1, int act1, void* o2, int ld2, int act2, cudaStream_t st);
// misc.cu
int pack_weight(int dtype, const float* w, int D0, int D1, void* p1, void* p2, cudaStream_t st);
int unpack_grad(const float* g, int D0, int D1, float* grad, int accumulate, cudaStream_t st);
int pack_input(int dtype, const float* s0, int c0, const float* s1, int c1, const float* s2, int c2,
               int N, int H, int W, int border, void* out, int Cpad, cudaStream_t st);
int pack_weight_thin(const float* w, int D0, int D1, int n_is_d0, int flip, void* out, cudaStream_t st);
int pack_weight_pad16(const float* w, int D0, int D1, int n_is_d0, void* out, cudaStream_t st);
int unpack_input_grad(int dtype, const void* g, int N, int H, int W, int ldg, int coff, int cn, float* grad,
                      int accumulate, cudaStream_t st);
int nhwc_to_nchw(int dtype, const void* x, int N, int H, int W, int C, int ld, float* out, cudaStream_t st);
int nchw_to_nhwc(int dtype, const float* x, int N, int H, int W, int C, void* out, int ld, cudaStream_t st);
int out_act_bwd(int dtype, int act, const float* o, const float* d, int N, int H, int W, int C, int border, void* g, int ldg,
                cudaStream_t st);
int fused_loss(const stcgan_loss_term* terms, int nterms, float* loss_out, cudaStream_t st);
int rel_logits(const float* a, const float* b, int N, long long M, int avg, int backward, float* out, cudaStream_t st);
int adam_step(const stcgan_adam_tensor* table, const int32_t* blocks, int nblocks, float* hyper, cudaStream_t st);
int adam_step_range(const stcgan_adam_tensor* table, const int32_t* blocks, int first_block, int nblocks, float* hyper, int tick,
                    int max_ctas, cudaStream_t st);
int float2uint(const float* in, long long n, uint8_t* out, cudaStream_t st);
int u8_to_nchw(const uint8_t* in, int N, int H, int W, int C, float* out, cudaStream_t st);
int float2uint_hwc(const float* in, int N, int C, int H, int W, uint8_t* out, cudaStream_t st);
}  // namespace stcgan

using namespace stcgan;

static bool dtype_ok(int d) { return d == STCGAN_F32 || d == STCGAN_BF16; }

extern "C" {

int stcgan_abi_version(void) { return STCGAN_ABI_VERSION; }
const char* stcgan_arch(void) { return "sm_100a"; }

const char* stcgan_error_string(int code) {
  if (code == 0) return "ok";
  if (code == STCGAN_EINVAL) return "stcgan: invalid argument (shape / alignment / enum)";
  if (code == STCGAN_EUNSUPPORTED) return "stcgan: shape not supported by the selected backend";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "stcgan: unknown error";
}

int64_t stcgan_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void stcgan_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

int stcgan_tapconv(int geom, int dtype, int backend, const void* x, int N, int IH, int IW, int K, int ldx,
                   const void* wp, const float* bias, int act, void* y, int OH, int OW, int Nout, int ldy,
                   int out_nchw_f32, void* workspace, int64_t workspace_bytes, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && x && wp && y);
  STCGAN_REQUIRE(N >= 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && K > 0 && Nout > 0 && ldx >= K);
  STCGAN_REQUIRE(out_nchw_f32 || ldy >= Nout);
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_SIGMOID);
  Geom g;
  if (!make_geom(geom, N, IH, IW, OH, OW, &g)) return STCGAN_EINVAL;
  if (N == 0) return 0;
  if (backend == STCGAN_BACKEND_TC) {
    if (dtype != STCGAN_BF16 || out_nchw_f32) return STCGAN_EUNSUPPORTED;
    return tapconv_tc(geom, g, x, K, ldx, wp, bias, act, y, Nout, ldy, as_stream(stream), 0, nullptr,
                      static_cast<float*>(workspace), (long long)workspace_bytes, nullptr, nullptr);
  }
  if (backend != STCGAN_BACKEND_FFMA) return STCGAN_EINVAL;
  return tapconv_ffma(g, dtype, x, K, ldx, wp, bias, act, y, Nout, ldy, out_nchw_f32, as_stream(stream));
}

int stcgan_tapconv_bnstats(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp,
                           void* y, int OH, int OW, int Nout, int ldy, void* workspace, int64_t workspace_bytes,
                           double* bn_acc, void* stream) {
  STCGAN_REQUIRE(x && wp && y && bn_acc);
  STCGAN_REQUIRE(N >= 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && K > 0 && Nout > 0 && ldx >= K && ldy >= Nout);
  Geom g;
  if (!make_geom(geom, N, IH, IW, OH, OW, &g)) return STCGAN_EINVAL;
  if (N == 0) return 0;
  return tapconv_tc(geom, g, x, K, ldx, wp, nullptr, STCGAN_ACT_NONE, y, Nout, ldy, as_stream(stream), 0, nullptr,
                    static_cast<float*>(workspace), (long long)workspace_bytes, bn_acc, nullptr);
}

int stcgan_tapconv_ep(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp,
                      const float* scale, const float* shift, int act, void* y, int ldy, int act2, void* y2, int ldy2,
                      int OH, int OW, int HC, int WC, int Nout, void* workspace, int64_t workspace_bytes, void* stream) {
  STCGAN_REQUIRE(x && wp && y);
  STCGAN_REQUIRE(N >= 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && K > 0 && Nout > 0 && ldx >= K && ldy >= Nout);
  STCGAN_REQUIRE(HC > 0 && WC > 0 && HC <= OH && WC <= OW && (!y2 || ldy2 >= Nout));
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_RELU && act2 >= STCGAN_ACT_NONE && act2 <= STCGAN_ACT_RELU);
  Geom g;
  if (!make_geom(geom, N, IH, IW, OH, OW, &g)) return STCGAN_EINVAL;
  if (N == 0) return 0;
  EpilogueExtra ex;
  ex.scale = scale; ex.y2 = y2; ex.ldy2 = ldy2; ex.act2 = act2; ex.HC = HC; ex.WC = WC;
  return tapconv_tc(geom, g, x, K, ldx, wp, shift, act, y, Nout, ldy, as_stream(stream), 0, nullptr,
                    static_cast<float*>(workspace), (long long)workspace_bytes, nullptr, &ex);
}

int stcgan_bn_fused_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy,
                          const double* acc, int64_t count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, int training,
                          float* mean_invstd, float* scale_shift, int HC, int WC,
                          void* out1, int ld1, int act1, void* out2, int ld2, int act2, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && N >= 0 && H > 0 && W > 0 && HC > 0 && WC > 0 && C > 0);
  return bn_fused_apply(dtype, y, N, H, W, C, ldy, acc, (long long)count, gamma, beta, running_mean, running_var, momentum,
                        eps, training, mean_invstd, scale_shift, HC, WC, out1, ld1, act1, out2, ld2, act2, as_stream(stream));
}

int stcgan_tapwgrad(int geom, int dtype, int backend, const void* S, int N, int SH, int SW, int D0, int lds,
                    const void* L, int LH, int LW, int D1, int ldl, float* G, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && S && L && G);
  STCGAN_REQUIRE(geom == STCGAN_GEOM_WIN_S2 || geom == STCGAN_GEOM_WIN_S1);
  STCGAN_REQUIRE(N >= 0 && SH > 0 && SW > 0 && LH > 0 && LW > 0 && D0 > 0 && D1 > 0 && lds >= D0 && ldl >= D1);
  if (N == 0) return 0;
  if (backend == STCGAN_BACKEND_TC) {
    if (dtype != STCGAN_BF16) return STCGAN_EUNSUPPORTED;
    return tapwgrad_tc(geom, S, N, SH, SW, D0, lds, L, LH, LW, D1, ldl, G, as_stream(stream));
  }
  if (backend != STCGAN_BACKEND_FFMA) return STCGAN_EINVAL;
  return tapwgrad_ffma(geom, dtype, S, N, SH, SW, D0, lds, L, LH, LW, D1, ldl, G, as_stream(stream));
}

int stcgan_pack_weight(int dtype, const float* w, int D0, int D1, void* p1, void* p2, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && w && (p1 || p2));
  return pack_weight(dtype, w, D0, D1, p1, p2, as_stream(stream));
}

int stcgan_unpack_grad(const float* g, int D0, int D1, float* grad, int accumulate, void* stream) {
  STCGAN_REQUIRE(g && grad);
  return unpack_grad(g, D0, D1, grad, accumulate, as_stream(stream));
}

int stcgan_bn_stats(int dtype, const void* y, int64_t P, int C, int ld, double* acc, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && acc && P >= 0 && C > 0 && ld >= C);
  return bn_stats(dtype, y, P, C, ld, acc, as_stream(stream));
}

int stcgan_bn_finalize(const double* acc, int64_t P, int C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int training,
                       float* mean_invstd, float* scale_shift, void* stream) {
  STCGAN_REQUIRE(C > 0 && gamma && beta && mean_invstd && scale_shift && (acc || !training) && (P > 0 || !training));
  return bn_finalize(acc, P, C, gamma, beta, running_mean, running_var, momentum, eps, training, mean_invstd,
                     scale_shift, as_stream(stream));
}

int stcgan_bn_running_update(const double* acc, int64_t count, int C, float* running_mean, float* running_var, float momentum,
                             void* stream) {
  STCGAN_REQUIRE(acc && count > 0 && C > 0 && running_mean && running_var);
  return bn_running_update(acc, count, C, running_mean, running_var, momentum, as_stream(stream));
}

int stcgan_bn_act_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* scale_shift,
                        int HC, int WC, void* out1, int ld1, int act1, void* out2, int ld2, int act2, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && N >= 0 && H > 0 && W > 0 && HC > 0 && WC > 0);
  return bn_act_apply(dtype, y, N, H, W, C, ldy, scale_shift, HC, WC, out1, ld1, act1, out2, ld2, act2, as_stream(stream));
}

int stcgan_bn_act_bwd_small(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* scale_shift,
                            const float* mean_invstd, const float* gamma, int HC, int WC, const void* g1, int ldg1, int act1,
                            const void* g2, int ldg2, int act2, void* dy, int lddy, float* dgamma, float* dbeta, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && N >= 0 && H > 0 && W > 0 && HC > 0 && WC > 0 && C > 0);
  if (N == 0) return 0;
  return bn_act_bwd_small(dtype, y, N, H, W, C, ldy, scale_shift, mean_invstd, gamma, HC, WC, g1, ldg1, act1, g2, ldg2, act2,
                          dy, lddy, dgamma, dbeta, as_stream(stream));
}

int stcgan_bn_act_bwd_reduce(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* scale_shift,
                             const float* mean_invstd, int HC, int WC, const void* g1, int ldg1, int act1,
                             const void* g2, int ldg2, int act2, double* acc, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && acc && HC <= H && WC <= W);
  return bn_act_bwd_reduce(dtype, y, N, H, W, C, ldy, scale_shift, mean_invstd, HC, WC, g1, ldg1, act1, g2, ldg2, act2,
                           acc, as_stream(stream));
}

int stcgan_bn_act_bwd_apply(int dtype, const void* y, int N, int H, int W, int C, int ldy, const float* scale_shift,
                            const float* mean_invstd, const float* gamma, int training, int HC, int WC,
                            const void* g1, int ldg1, int act1, const void* g2, int ldg2, int act2,
                            const double* acc, void* dy, int lddy, float* dgamma, float* dbeta, float* dbias,
                            void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && y && HC <= H && WC <= W);
  return bn_act_bwd_apply(dtype, y, N, H, W, C, ldy, scale_shift, mean_invstd, gamma, training, HC, WC, g1, ldg1, act1,
                          g2, ldg2, act2, acc, dy, lddy, dgamma, dbeta, dbias, as_stream(stream));
}

int stcgan_colsum(int dtype, const void* g, int64_t P, int C, int ld, float* out, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && g && out && P >= 0 && C > 0 && ld >= C);
  return colsum(dtype, g, P, C, ld, out, as_stream(stream));
}

int stcgan_pack_input(int dtype, const float* s0, int c0, const float* s1, int c1, const float* s2, int c2,
                      int N, int H, int W, int border, void* out, int Cpad, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && out && (c0 == 0 || s0) && (c1 == 0 || s1) && (c2 == 0 || s2));
  return pack_input(dtype, s0, c0, s1, c1, s2, c2, N, H, W, border, out, Cpad, as_stream(stream));
}

int stcgan_tapconv_thin_n(int geom, const void* x, int N, int IH, int IW, int K, int ldx, const void* wp16,
                          const float* bias, int act, void* y_nhwc8, int ldy, float* y_nchw_f32, int OH, int OW, int Nout,
                          void* stream) {
  STCGAN_REQUIRE(x && wp16 && (y_nhwc8 || y_nchw_f32) && N >= 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && K > 0 && ldx >= K);
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_SIGMOID);
  Geom g;
  if (!make_geom(geom, N, IH, IW, OH, OW, &g)) return STCGAN_EINVAL;
  if (N == 0) return 0;
  return tapconv_tc(geom, g, x, K, ldx, wp16, bias, act, y_nhwc8, Nout, ldy, as_stream(stream), 1, y_nchw_f32, nullptr, 0, nullptr,
                    nullptr);
}

int stcgan_thinconv(const void* t, int N, int HP, int WP, int stride, const void* wthin, const float* bias, int act,
                    void* y, int OH, int OW, int Nout, int ldy, void* stream) {
  STCGAN_REQUIRE(t && wthin && y && N >= 0 && HP >= 4 && WP >= 4 && (stride == 1 || stride == 2) && OH > 0 && OW > 0 && ldy >= Nout);
  if (N == 0) return 0;
  return thinconv_tc(t, N, HP, WP, stride, wthin, bias, act, y, OH, OW, Nout, ldy, as_stream(stream), nullptr);
}

int stcgan_thinconv2(const void* t, int N, int HP, int WP, int stride, const void* wthin, const float* bias, int act,
                     void* y, int ldy, int act2, void* y2, int ldy2, int OH, int OW, int Nout, void* stream) {
  STCGAN_REQUIRE(t && wthin && y && y2 && N >= 0 && HP > 0 && WP > 0 && (stride == 1 || stride == 2));
  STCGAN_REQUIRE(OH > 0 && OW > 0 && Nout > 0 && ldy >= Nout && ldy2 >= Nout);
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_RELU && act2 >= STCGAN_ACT_NONE && act2 <= STCGAN_ACT_RELU);
  if (N == 0) return 0;
  EpilogueExtra ex;
  ex.y2 = y2; ex.ldy2 = ldy2; ex.act2 = act2;
  return thinconv_tc(t, N, HP, WP, stride, wthin, bias, act, y, OH, OW, Nout, ldy, as_stream(stream), &ex);
}

int stcgan_thinwgrad(const void* t, int N, int HP, int WP, int stride, int thin_c, const void* f, int FH, int FW, int Dfat,
                     int ldf, int fat_is_dim0, int flip, float* G, void* stream) {
  STCGAN_REQUIRE(t && f && G && N >= 0 && HP >= 4 && WP >= 4 && (stride == 1 || stride == 2) && FH > 0 && FW > 0 && ldf >= Dfat);
  if (N == 0) return 0;
  return thinwgrad_tc(t, N, HP, WP, stride, thin_c, f, FH, FW, Dfat, ldf, fat_is_dim0, flip, G, as_stream(stream));
}

int stcgan_thin_col2im(int mode, const void* x, int N, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                       const float* bias, int act, float* y_nchw_f32, void* y_nhwc8, int ldy, int OH, int OW, void* stream) {
  STCGAN_REQUIRE(x && wt && N >= 0 && IH > 0 && IW > 0 && K > 0 && ldx >= K && OH > 0 && OW > 0);
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_SIGMOID);
  STCGAN_REQUIRE(mode == 0 ? (OH <= 2 * IH + 1 && OW <= 2 * IW + 1) : (mode == 1 && OH == IH - 1 && OW == IW - 1));
  if (N == 0) return 0;
  return thin_col2im_tc(mode, x, N, IH, IW, K, ldx, wt, cpad, cout, bias, act, y_nchw_f32, y_nhwc8, ldy, OH, OW, as_stream(stream),
                        nullptr);
}

int stcgan_thin_convt_u8(const void* x, int N, int IH, int IW, int K, int ldx, const void* wt, int cpad, int cout,
                         const float* bias, int act, float* y_nchw_f32, uint8_t* y_nhwc_u8, int OH, int OW, void* stream) {
  STCGAN_REQUIRE(x && wt && y_nhwc_u8 && N >= 0 && IH > 0 && IW > 0 && K > 0 && ldx >= K && OH > 0 && OW > 0);
  STCGAN_REQUIRE(act >= STCGAN_ACT_NONE && act <= STCGAN_ACT_SIGMOID && OH <= 2 * IH + 1 && OW <= 2 * IW + 1);
  if (N == 0) return 0;
  return thin_col2im_tc(0, x, N, IH, IW, K, ldx, wt, cpad, cout, bias, act, y_nchw_f32, nullptr, 0, OH, OW, as_stream(stream),
                        y_nhwc_u8);
}

int stcgan_pack_weight_tapn(const float* w, int D0, int D1, int n_is_d0, int cpad, void* out, void* stream) {
  STCGAN_REQUIRE(w && out && D0 > 0 && D1 > 0);
  return pack_weight_tapn(w, D0, D1, n_is_d0, cpad, out, as_stream(stream));
}

int stcgan_pack_weight_thin(const float* w, int D0, int D1, int n_is_d0, int flip, void* out, void* stream) {
  STCGAN_REQUIRE(w && out && D0 > 0 && D1 > 0);
  return pack_weight_thin(w, D0, D1, n_is_d0, flip, out, as_stream(stream));
}

int stcgan_pack_weight_pad16(const float* w, int D0, int D1, int n_is_d0, void* out, void* stream) {
  STCGAN_REQUIRE(w && out && D0 > 0 && D1 > 0);
  return pack_weight_pad16(w, D0, D1, n_is_d0, out, as_stream(stream));
}

int stcgan_unpack_input_grad(int dtype, const void* g, int N, int H, int W, int ldg, int coff, int cn,
                             float* grad_nchw, int accumulate, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && g && grad_nchw && coff >= 0 && cn >= 0 && coff + cn <= ldg);
  return unpack_input_grad(dtype, g, N, H, W, ldg, coff, cn, grad_nchw, accumulate, as_stream(stream));
}

int stcgan_nhwc_to_nchw(int dtype, const void* x, int N, int H, int W, int C, int ld, float* out, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && x && out && ld >= C);
  return nhwc_to_nchw(dtype, x, N, H, W, C, ld, out, as_stream(stream));
}

int stcgan_nchw_to_nhwc(int dtype, const float* x, int N, int H, int W, int C, void* out, int ld, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && x && out && ld >= C);
  return nchw_to_nhwc(dtype, x, N, H, W, C, out, ld, as_stream(stream));
}

int stcgan_out_act_bwd(int dtype, int act, const float* out_nchw, const float* dout_nchw, int N, int H, int W, int C,
                       int border, void* g, int ldg, void* stream) {
  STCGAN_REQUIRE(dtype_ok(dtype) && out_nchw && dout_nchw && g && ldg >= C && border >= 0);
  return out_act_bwd(dtype, act, out_nchw, dout_nchw, N, H, W, C, border, g, ldg, as_stream(stream));
}

int stcgan_fused_loss(const stcgan_loss_term* host_terms, int nterms, float* loss_out, void* stream) {
  STCGAN_REQUIRE(host_terms && loss_out);
  return fused_loss(host_terms, nterms, loss_out, as_stream(stream));
}

int stcgan_rel_logits(const float* a, const float* b, int N, int64_t M, int avg, int backward, float* out, void* stream) {
  STCGAN_REQUIRE(b && out && (a || backward) && N >= 0 && M >= 0);
  return rel_logits(a, b, N, (long long)M, avg, backward, out, as_stream(stream));
}

int stcgan_adam_step(const stcgan_adam_tensor* dev_table, const int32_t* dev_blocks, int nblocks,
                     float* dev_hyper, void* stream) {
  STCGAN_REQUIRE(dev_table && dev_blocks && dev_hyper);
  return adam_step(dev_table, dev_blocks, nblocks, dev_hyper, as_stream(stream));
}

int stcgan_adam_step_range(const stcgan_adam_tensor* dev_table, const int32_t* dev_blocks, int first_block, int nblocks,
                           float* dev_hyper, int tick, int max_ctas, void* stream) {
  STCGAN_REQUIRE(dev_table && dev_blocks && dev_hyper);
  return adam_step_range(dev_table, dev_blocks, first_block, nblocks, dev_hyper, tick, max_ctas, as_stream(stream));
}

int stcgan_adam_chunk(void) { return 256 * 16; }
int stcgan_adam_tile(void) { return 32; }

int stcgan_float2uint_hwc(const float* nchw, int N, int C, int H, int W, uint8_t* out_nhwc, void* stream) {
  STCGAN_REQUIRE(nchw && out_nhwc);
  return float2uint_hwc(nchw, N, C, H, W, out_nhwc, as_stream(stream));
}

int stcgan_augment_u8(const uint8_t* img_nhwc, int N, int H, int W, int C, const stcgan_aug_sample* dev_samples,
                      int crop_h, int crop_w, float* out_nchw, void* stream) {
  STCGAN_REQUIRE(img_nhwc && dev_samples && out_nchw && N >= 0 && H > 0 && W > 0 && crop_h > 0 && crop_w > 0);
  return augment_u8(img_nhwc, N, H, W, C, dev_samples, crop_h, crop_w, out_nchw, as_stream(stream));
}

int stcgan_u8_hwc_to_nchw_f32(const uint8_t* in_nhwc, int N, int H, int W, int C, float* out_nchw, void* stream) {
  STCGAN_REQUIRE(in_nhwc && out_nchw && N >= 0 && H > 0 && W > 0 && C > 0);
  return u8_to_nchw(in_nhwc, N, H, W, C, out_nchw, as_stream(stream));
}

int stcgan_float2uint(const float* in, int64_t n, uint8_t* out, void* stream) {
  STCGAN_REQUIRE(in && out && n >= 0);
  return float2uint(in, n, out, as_stream(stream));
}

}  // extern "C"

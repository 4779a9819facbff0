// Training augmentation + dataset transform on the GPU.
//
// Reference: src/transform.py:57-156 composed by src/cgan.py:105-110 (RandomScale -> RandomRotate -> RandomHorizontalFlip ->
// RandomCrop on float32 HWC images in [0,1]) after utils.uint2float (src/utils.py:60-62) and before the dataset's
// (s.transpose(2,0,1) - 0.5) * 2 (src/dataset.py:152).  The reference runs this with OpenCV on DataLoader worker processes;
// at > 2.5 k images/s per GPU that loader is the bottleneck of a real run, so here the host only draws the five random
// numbers per sample (in the reference's order) and ships the decoded uint8 image.
//
// Both warps are cv.warpAffine(x, getRotationMatrix2D(centre, angle, scale), (cols, rows), INTER_LINEAR, BORDER_CONSTANT 0)
// (INTER_AREA is INTER_LINEAR inside warpAffine) and are reproduced BIT FOR BIT: OpenCV forms the source position of a
// destination pixel in fixed point (10 fractional bits, rounded to 1/32 pixel), weights the four taps (0 outside the image)
// with float32 products of (1 - f) and f and sums them left to right -- integer arithmetic and a fixed float32 expression,
// so the GPU result equals cv2's exactly.  The warps are applied one after the other (the second resamples the first
// one's OUTPUT, as the reference does), per output pixel: 4 taps of the scaled image, each 4 taps of the uint8 source.  The
// host passes the two INVERSE matrices in float64 (formed exactly like cv::warpAffine forms them).
#include "common.cuh"

namespace stcgan {

struct AugSample {          // one per image of the batch
  double si[6];             // inverse of the RandomScale matrix   (row-major 2 x 3)
  double ri[6];             // inverse of the RandomRotate matrix
  int flip, row_off, col_off, identity;   // identity != 0: scale == 1 and angle == 0 (both warps are exact copies)
};

template <int C>
__device__ __forceinline__ void src_px(const uint8_t* __restrict__ img, int H, int W, int x, int y, float* v) {
  if (x < 0 || y < 0 || x >= W || y >= H) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = 0.f;
    return;
  }
  const uint8_t* p = img + ((long long)y * W + x) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = __fdiv_rn((float)p[c], 255.f);          // utils.uint2float
}

// cv::warpAffine's source position for destination pixel (x, y): fixed point with AB_BITS = 10, rounded to 1/32 pixel
// (INTER_BITS = 5): returns the integer tap (sx, sy) and the fractions (fx, fy) in 1/32 units.  Products / sums are formed
// without fused multiply-add, as the (x86 baseline) OpenCV build forms them; saturate_cast<int> = round half to even.
__device__ __forceinline__ void warp_pos(const double* m, int x, int y, int* sx, int* sy, int* fx, int* fy) {
  const long long adelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
  const long long bdelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
  const long long X0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
  const long long Y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
  const long long X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  *sx = (int)(X >> 5); *sy = (int)(Y >> 5); *fx = (int)(X & 31); *fy = (int)(Y & 31);
}

// cv::remap's float32 bilinear weights (BilinearTab_f) and left-to-right sum
template <int C>
__device__ __forceinline__ void blend(const float* a, const float* b, const float* c2, const float* d, int fxi, int fyi, float* out) {
  const float fx = __fdiv_rn((float)fxi, 32.f), fy = __fdiv_rn((float)fyi, 32.f);
  const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
  const float w00 = __fmul_rn(gy, gx), w01 = __fmul_rn(gy, fx), w10 = __fmul_rn(fy, gx), w11 = __fmul_rn(fy, fx);
#pragma unroll
  for (int c = 0; c < C; ++c)
    out[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[c], w00), __fmul_rn(b[c], w01)), __fmul_rn(c2[c], w10)), __fmul_rn(d[c], w11));
}

// value of the SCALED image at the integer position (u, v)
template <int C>
__device__ __forceinline__ void scaled_px(const uint8_t* __restrict__ img, int H, int W, const double* si, int u, int v, float* out) {
  if (u < 0 || v < 0 || u >= W || v >= H) {
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = 0.f;
    return;
  }
  int x0, y0, fx, fy;
  warp_pos(si, u, v, &x0, &y0, &fx, &fy);
  float a[C], b[C], c2[C], d[C];
  src_px<C>(img, H, W, x0, y0, a); src_px<C>(img, H, W, x0 + 1, y0, b);
  src_px<C>(img, H, W, x0, y0 + 1, c2); src_px<C>(img, H, W, x0 + 1, y0 + 1, d);
  blend<C>(a, b, c2, d, fx, fy, out);
}

template <int C>
__global__ void __launch_bounds__(256)
augment_kernel(const uint8_t* __restrict__ img, int N, int H, int W, const AugSample* __restrict__ samples, int CH, int CW,
               float* __restrict__ out) {
  pdl_prologue();
  const long long total = (long long)N * CH * CW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % CW); const long long t = i / CW;
    const int r = (int)(t % CH); const int n = (int)(t / CH);
    const AugSample s = samples[n];
    const uint8_t* im = img + (long long)n * H * W * C;
    const int y = r + s.row_off;                                   // RandomCrop (transform.py:137-148)
    int x = j + s.col_off;
    if (s.flip) x = W - 1 - x;                                     // np.fliplr before the crop (transform.py:111)
    float v[C];
    if (s.identity) {
      src_px<C>(im, H, W, x, y, v);
    } else {
      int u0, v0, fx, fy;
      warp_pos(s.ri, x, y, &u0, &v0, &fx, &fy);
      float a[C], b[C], c2[C], d[C];
      scaled_px<C>(im, H, W, s.si, u0, v0, a); scaled_px<C>(im, H, W, s.si, u0 + 1, v0, b);
      scaled_px<C>(im, H, W, s.si, u0, v0 + 1, c2); scaled_px<C>(im, H, W, s.si, u0 + 1, v0 + 1, d);
      blend<C>(a, b, c2, d, fx, fy, v);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)                                    // (s.transpose(2,0,1) - 0.5) * 2   (dataset.py:152)
      out[(((long long)n * C + c) * CH + r) * CW + j] = __fmul_rn(__fsub_rn(v[c], 0.5f), 2.f);
  }
}

int augment_u8(const uint8_t* img, int N, int H, int W, int C, const void* samples, int CH, int CW, float* out, cudaStream_t st) {
  if (C != 1 && C != 3) return STCGAN_EUNSUPPORTED;
  if (CH > H || CW > W) return STCGAN_EUNSUPPORTED;      // (RandomCrop's zero-padding branch: not reachable for ISTD, 640x480 >= 256)
  const long long total = (long long)N * CH * CW;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
  const AugSample* s = static_cast<const AugSample*>(samples);
  if (C == 3) launch_k(augment_kernel<3>, (unsigned)blocks, 256, 0, st, img, N, H, W, s, CH, CW, out);
  else launch_k(augment_kernel<1>, (unsigned)blocks, 256, 0, st, img, N, H, W, s, CH, CW, out);
  return finish_launch();
}

}  // namespace stcgan

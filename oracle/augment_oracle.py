"""CPU restatement of the reference's training augmentation + dataset transform.  TEST INFRASTRUCTURE ONLY (see
stcgan_oracle.py's header; imported by tests/ only).

Follows src/transform.py:57-156 (RandomScale, RandomRotate, RandomHorizontalFlip, RandomCrop as composed by
src/cgan.py:105-110), src/utils.py:60-62 (uint2float) and src/dataset.py:152.  The two cv.warpAffine calls belong to a
third-party dependency that is not vendored (OpenCV, pinned opencv 4.4.0 in environment.yml; this container has 4.13): its
published algorithm for float32 / INTER_LINEAR / BORDER_CONSTANT is restated in `warp_bilinear` (fixed-point source
positions at 1/32 pixel, float32 tap weights; INTER_AREA is INTER_LINEAR inside warpAffine).  Pinning: bit-for-bit against
`cv2.warpAffine` here, and against `tests/golden/augment_vectors.npz`, produced by running the reference's own transform
classes on this container's OpenCV (tests/golden/make_golden_augment.py).
"""
import math

import numpy as np


def rotation_matrix(cx, cy, angle_deg, scale):
    a = angle_deg * math.pi / 180.0
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def invert_affine(m):
    """the inversion cv::warpAffine applies to M when WARP_INVERSE_MAP is not set (float64)"""
    m = np.asarray(m, dtype=np.float64).reshape(-1)
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    i1, i3 = m[1] * -d, m[3] * -d
    b1 = -a11 * m[2] - i1 * m[5]
    b2 = -i3 * m[2] - a22 * m[5]
    return np.array([[a11, i1, b1], [i3, a22, b2]], dtype=np.float64)


def warp_bilinear(src, minv):
    """cv::warpAffine + cv::remap for float32 images, INTER_LINEAR, BORDER_CONSTANT(0), as OpenCV computes it: source
    positions in FIXED POINT (AB_BITS = 10, rounded to 1/32 pixel = INTER_BITS 5), the four taps weighted with float32
    products (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx and summed left to right.  Checked bit-for-bit against
    cv2.warpAffine 4.13 (tests/test_augment_cpu.py)."""
    h, w, c = src.shape
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta = np.rint(minv[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(minv[1, 0] * xs * 1024).astype(np.int64)
    x0_ = np.rint((minv[0, 1] * ys + minv[0, 2]) * 1024).astype(np.int64) + 16
    y0_ = np.rint((minv[1, 1] * ys + minv[1, 2]) * 1024).astype(np.int64) + 16
    X, Y = (x0_[:, None] + adelta[None, :]) >> 5, (y0_[:, None] + bdelta[None, :]) >> 5
    x0, y0 = X >> 5, Y >> 5
    fx = ((X & 31).astype(np.float32) / np.float32(32))[..., None]
    fy = ((Y & 31).astype(np.float32) / np.float32(32))[..., None]

    def tap(xi, yi):
        ok = (xi >= 0) & (yi >= 0) & (xi < w) & (yi < h)
        v = src[np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)]
        return np.where(ok[..., None], v, np.float32(0))

    one = np.float32(1)
    return (tap(x0, y0) * ((one - fy) * (one - fx)) + tap(x0 + 1, y0) * ((one - fy) * fx)
            + tap(x0, y0 + 1) * (fy * (one - fx)) + tap(x0 + 1, y0 + 1) * (fy * fx)).astype(np.float32)


def augment(img_u8, p, crop):
    """img_u8 uint8 [H,W,C]; p = dict(scale, angle, flip, row_off, col_off) -> float32 [C, crop, crop] in [-1, 1]"""
    x = img_u8.astype(np.float32) / 255                                   # utils.uint2float
    h, w = x.shape[:2]
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    if not (p["scale"] == 1.0 and p["angle"] == 0.0):
        x = warp_bilinear(x, invert_affine(rotation_matrix(cx, cy, 0.0, p["scale"])))   # RandomScale (transform.py:57-75)
        x = warp_bilinear(x, invert_affine(rotation_matrix(cx, cy, p["angle"], 1.0)))   # RandomRotate (transform.py:78-100)
    if p["flip"]:
        x = np.fliplr(x)                                                  # transform.py:103-117
    x = x[p["row_off"]:p["row_off"] + crop, p["col_off"]:p["col_off"] + crop]          # transform.py:120-156
    return ((x.transpose(2, 0, 1) - np.float32(0.5)) * np.float32(2)).astype(np.float32)   # dataset.py:152

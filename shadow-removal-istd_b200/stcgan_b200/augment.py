"""Host side of the GPU input pipeline: the random draws of the reference's training augmentation, in the
reference's order, and the launch of the fused augmentation + dataset-transform kernel.

Reference: `transform.transforms(scale=, angle=, flip_prob=0.5, crop_size=)` (src/cgan.py:105-110) = Compose([RandomScale,
RandomRotate, RandomHorizontalFlip, RandomCrop]) (src/transform.py:7-25), applied to the float images of one sample with
ONE set of random numbers per sample (src/transform.py:57-156), after utils.uint2float and before `(s - 0.5) * 2` + HWC->CHW
(src/dataset.py:100-110, 152).  Everything numeric runs in `stcgan_augment_u8`; this module draws
    scale ~ U(1 - s, 1 + s);  angle ~ U(-a, a);  flip = not (rand() > p);  row, col = randint(0, rows - crop), randint(0, cols - crop)
with the caller's `numpy.random` generator -- the same calls in the same order as the reference, so a loader that seeds like
the reference's workers (np.random.seed(42 + id), src/cgan.py:123-124) reproduces its augmentation stream.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib, ops


def _rotation_matrix(cx, cy, angle_deg, scale):
    """cv.getRotationMatrix2D((cx, cy), angle, scale) in float64"""
    a = angle_deg * math.pi / 180.0
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    return [alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy]


def _invert_affine(m):
    """the inversion cv::warpAffine applies to M (no WARP_INVERSE_MAP), float64"""
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    i0, i1, i3, i4 = a11, m[1] * -d, m[3] * -d, a22
    b1 = -i0 * m[2] - i1 * m[5]
    b2 = -i3 * m[2] - i4 * m[5]
    return [i0, i1, b1, i3, i4, b2]


def sample_params(rng, n, height, width, scale=0.05, angle=15, flip_prob=0.5, crop=256):
    """Draw the augmentation of `n` samples with `rng` (the `numpy.random` module or a RandomState) in the reference's call
    order.  `scale` / `angle` / `flip_prob` None = that transform is absent (transform.transforms skips it).  Returns a list of
    dicts (scale, angle, flip, row_off, col_off)."""
    ch, cw = (crop, crop) if isinstance(crop, int) else crop
    out = []
    for _ in range(n):
        s = float(rng.uniform(low=1.0 - scale, high=1.0 + scale)) if scale is not None else 1.0       # transform.py:64
        a = float(rng.uniform(low=-angle, high=angle)) if angle is not None else 0.0                   # transform.py:89
        f = (not (rng.rand() > flip_prob)) if flip_prob is not None else False                         # transform.py:108
        ro = int(rng.randint(low=0, high=height - ch))                                                  # transform.py:137
        co = int(rng.randint(low=0, high=width - cw))                                                   # transform.py:138
        out.append(dict(scale=s, angle=a, flip=bool(f), row_off=ro, col_off=co))
    return out


def pack_params(params, height, width, device):
    """list of sample dicts -> device table of stcgan_aug_sample"""
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    arr = (_lib.AugSample * len(params))()
    for i, p in enumerate(params):
        si = _invert_affine(_rotation_matrix(cx, cy, 0.0, p["scale"]))
        ri = _invert_affine(_rotation_matrix(cx, cy, p["angle"], 1.0))
        arr[i].scale_inv[:] = si
        arr[i].rot_inv[:] = ri
        arr[i].flip, arr[i].row_off, arr[i].col_off = int(p["flip"]), int(p["row_off"]), int(p["col_off"])
        arr[i].identity = int(p["scale"] == 1.0 and p["angle"] == 0.0)
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def augment_u8(img_u8, params, crop, out=None):
    """img_u8: uint8 CUDA tensor [N,H,W,C] (C = 1 or 3) as cv2.imread gives it; `params`: sample_params(...) output (the SAME
    list for the image, the matte and the target of a batch) or an already packed table.  Returns float32 [N,C,crop,crop]."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or not img_u8.is_contiguous():
        raise ValueError("augment_u8 expects a contiguous uint8 CUDA tensor [N,H,W,C]")
    n, h, w, c = img_u8.shape
    ch, cw = (crop, crop) if isinstance(crop, int) else crop
    table = params if isinstance(params, torch.Tensor) else pack_params(params, h, w, img_u8.device)
    if table.numel() != n * C.sizeof(_lib.AugSample):
        raise ValueError("one parameter record per image expected")
    if out is None:
        out = torch.empty((n, c, ch, cw), dtype=torch.float32, device=img_u8.device)
    _lib.check(_lib.load().stcgan_augment_u8(img_u8.data_ptr(), n, h, w, c, table.data_ptr(), ch, cw, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "stcgan_augment_u8")
    return out

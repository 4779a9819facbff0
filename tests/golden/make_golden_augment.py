"""Golden vectors for the augmentation path: the REFERENCE's own transform classes (src/transform.py, imported unmodified from
/root/reference) on this container's OpenCV, with the reference's worker seeding (np.random.seed(42 + id), src/cgan.py:123).

    python tests/golden/make_golden_augment.py      ->  tests/golden/augment_vectors.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import cv2  # noqa: E402,F401
from src import transform, utils  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
H, W, CROP, N = 60, 84, 32, 6


def smooth_u8(rng, c):
    lo = rng.rand(H // 6, W // 6, c).astype(np.float32)
    z = np.repeat(np.repeat(lo, 6, 0), 6, 1)[:H, :W]
    z = cv2.GaussianBlur(z, (0, 0), 2.0).reshape(H, W, c)
    return np.clip(z * 255, 0, 255).astype(np.uint8)


def main():
    out = {}
    for case, kw in (("full", dict(scale=0.05, angle=15, flip_prob=0.5, crop_size=CROP)),
                     ("flipcrop", dict(flip_prob=0.5, crop_size=CROP))):
        rng = np.random.RandomState(7)
        imgs = np.stack([smooth_u8(rng, 3) for _ in range(N)])
        mattes = np.stack([smooth_u8(rng, 1) for _ in range(N)])
        tf = transform.transforms(**kw)
        np.random.seed(42)                                          # worker 0 of the reference's DataLoader
        o_img, o_mat = [], []
        for i in range(N):
            a, b = tf(utils.uint2float(imgs[i]), utils.uint2float(mattes[i][..., 0]))
            if b.ndim == 2:
                b = b[:, :, np.newaxis]                             # dataset.py:141-143
            o_img.append(((a.transpose(2, 0, 1) - 0.5) * 2).astype(np.float32))       # dataset.py:152
            o_mat.append(((b.transpose(2, 0, 1) - 0.5) * 2).astype(np.float32))
        out.update({f"{case}/img_u8": imgs, f"{case}/matte_u8": mattes, f"{case}/img_out": np.stack(o_img),
                    f"{case}/matte_out": np.stack(o_mat)})
    out["meta"] = np.array([H, W, CROP, N])
    out["opencv_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "augment_vectors.npz"), **out)
    print("wrote augment_vectors.npz with OpenCV", cv2.__version__)


if __name__ == "__main__":
    main()

"""The augmentation path on the CPU: (1) stcgan_b200.augment.sample_params draws what the reference's transform classes draw
(same numpy.random calls in the same order, src/transform.py:64,89,108,137-138) -- checked by re-creating the reference's golden
outputs with the CPU restatement driven by OUR draws; (2) the restatement (oracle/augment_oracle.py) matches the reference +
OpenCV golden vectors (tests/golden/make_golden_augment.py) bit for bit."""
import os

import numpy as np
import pytest

from conftest import ROOT

import augment_oracle as AO


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(ROOT, "tests", "golden", "augment_vectors.npz"))


@pytest.mark.parametrize("case,kw", [("full", dict(scale=0.05, angle=15, flip_prob=0.5)),
                                     ("flipcrop", dict(scale=None, angle=None, flip_prob=0.5))])
def test_sampler_and_restatement_reproduce_reference_augmentation(vec, case, kw):
    from stcgan_b200 import augment as A
    h, w, crop, n = (int(v) for v in vec["meta"])
    np.random.seed(42)                                             # the reference's worker-0 seed (src/cgan.py:123-124)
    params = A.sample_params(np.random, n, h, w, crop=crop, **kw)
    assert any(p["flip"] for p in params) and not all(p["flip"] for p in params)
    worst = 0.0
    for i, p in enumerate(params):
        for key, c in (("img", 3), ("matte", 1)):
            got = AO.augment(vec[f"{case}/{key}_u8"][i], p, crop)
            want = vec[f"{case}/{key}_out"][i]
            assert got.shape == want.shape == (c, crop, crop)
            worst = max(worst, float(np.abs(got - want).max()))
    assert worst == 0.0, worst          # bit-exact: integer source positions + a fixed float32 expression (see warp_bilinear)


def test_affine_helpers_match_opencv():
    cv2 = pytest.importorskip("cv2")
    from stcgan_b200 import augment as A
    m = np.array(A._rotation_matrix(41.5, 29.5, 11.0, 1.0)).reshape(2, 3)
    assert np.allclose(m, cv2.getRotationMatrix2D((41.5, 29.5), 11.0, 1.0), rtol=0, atol=1e-12)
    inv = np.array(A._invert_affine(list(m.reshape(-1)))).reshape(2, 3)
    assert np.allclose(inv, cv2.invertAffineTransform(m), rtol=0, atol=1e-12)
    assert np.allclose(inv, AO.invert_affine(m), rtol=0, atol=1e-12)

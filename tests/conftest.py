import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "shadow-removal-istd_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand; building needs nvcc but no GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("stcgan_build", os.path.join(ROOT, "shadow-removal-istd_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.OUT):
        mod.build()
    from stcgan_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / ||b|| in float64 (CPU)."""
    import torch
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def golden():
    """Fixtures written by tests/golden/make_golden.py from the unmodified reference."""
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    return {
        "step": np.load(os.path.join(gdir, "stcgan_step_b2_256.npz")),
        "adv": np.load(os.path.join(gdir, "adversarial_loss.npz")),
        "infer": np.load(os.path.join(gdir, "stcgan_infer_480x640.npz")),
        "f2u": np.load(os.path.join(gdir, "float2uint_vectors.npz")),
        "weights": np.load(os.path.join(gdir, "weights_fingerprint.npz")),
    }

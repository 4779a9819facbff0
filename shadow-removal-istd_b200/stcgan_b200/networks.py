"""Model registry with the reference's interface (src/networks.py:19-54) for the ST-CGAN keys, plus the
shim that substitutes these classes into an importable copy of the reference so that `src/cgan.py` /
`src/main.py` build the B200 modules through their own `networks.get_generator("stcgan", ...)` calls.
"""
from __future__ import annotations

import sys
from enum import Enum, unique

import torch
import torch.nn as nn

from .modules import NLayerDiscriminator, UnetGenerator


@torch.no_grad()
def weights_init(m):
    """Same rule as src/networks.py:19-30: class-name match, N(0, 0.02) weights, zero bias
    (and N(1, 0.02) for 'Linear', which these networks do not contain)."""
    name = type(m).__name__
    if "Conv" in name or "BatchNorm" in name:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias.data, 0)
    elif "Linear" in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias.data, 0)


@unique
class Generators(Enum):
    STCGAN = UnetGenerator


@unique
class Discriminators(Enum):
    STCGAN = NLayerDiscriminator


def get_generator(key: str, *args, **kwargs):
    try:
        cls = Generators[key.upper()].value
    except KeyError:
        raise KeyError(f"stcgan_b200 provides the 'stcgan' generator only (got {key!r}); the reference's other "
                       "generators are outside the accelerated path") from None
    return cls(*args, **kwargs)


def get_discriminator(key: str, *args, **kwargs):
    try:
        cls = Discriminators[key.upper()].value
    except KeyError:
        raise KeyError(f"stcgan_b200 provides the 'stcgan' discriminator only (got {key!r})") from None
    return cls(*args, **kwargs)


def install_into_reference(reference_root=None):
    """Make the reference build B200 modules: call BEFORE `import src.networks`.

    Pre-imports `src.models.stcgan_g` / `src.models.stcgan_d` from the reference tree and rebinds their class
    names to this package's classes, so that `src/networks.py:33-46` picks them up when it builds its enums
    (the reference is read-only; nothing is written).  Also rebinds `src.loss.AdversarialLoss/DataLoss` when
    that module is already imported.  Returns the patched modules.
    """
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import importlib
    g = importlib.import_module("src.models.stcgan_g")
    d = importlib.import_module("src.models.stcgan_d")
    g.UnetGenerator = UnetGenerator
    d.NLayerDiscriminator = NLayerDiscriminator
    if "src.networks" in sys.modules:          # already imported: patch the registry members in place
        nw = sys.modules["src.networks"]
        nw.UnetGenerator, nw.NLayerDiscriminator = UnetGenerator, NLayerDiscriminator
        nw.get_generator = lambda key, *a, **k: (get_generator(key, *a, **k) if key.lower() == "stcgan"
                                                 else nw.Generators[key.upper()].value(*a, **k))
        nw.get_discriminator = lambda key, *a, **k: (get_discriminator(key, *a, **k) if key.lower() == "stcgan"
                                                     else nw.Discriminators[key.upper()].value(*a, **k))
    return g, d

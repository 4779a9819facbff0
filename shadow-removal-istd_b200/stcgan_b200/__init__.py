"""stcgan_b200 -- B200-native (sm_100a) implementation of the ST-CGAN hot path of nhchiu/Shadow-Removal-ISTD.

Public surface (mirrors the reference's src/networks.py, src/models/stcgan_{g,d}.py, src/loss.py):

    UnetGenerator, NLayerDiscriminator          drop-in nn.Modules (same ctor/forward/state_dict)
    get_generator, get_discriminator, weights_init, install_into_reference
    AdversarialLoss, DataLoss                   fused loss kernels behind the reference's loss API
    FusedAdam                                   torch.optim.Adam-compatible fused optimiser
    STCGANEngine, TrainConfig, infer, infer_u8  the hand-scheduled train step / inference (src/cgan.py:274-351, 437-446)
    InferencePipeline                           infer_u8 over a stream of host batches with the transfers overlapped

All arithmetic runs in libstcgan_b200.so (hand-written CUDA for sm_100a, C ABI in include/stcgan_b200.h).
There is no CPU or eager-PyTorch fallback; loading fails loudly if the library is missing.
"""
from . import _lib
from ._lib import StcganError, StcganLibraryError
from .engine import InferencePipeline, STCGANEngine, TrainConfig, infer, infer_u8
from .loss import AdversarialLoss, DataLoss
from .modules import NLayerDiscriminator, UnetGenerator
from .networks import (Discriminators, Generators, get_discriminator, get_generator, install_into_reference,
                       weights_init)
from .optim import FusedAdam

__all__ = ["UnetGenerator", "NLayerDiscriminator", "get_generator", "get_discriminator", "weights_init",
           "install_into_reference", "AdversarialLoss", "DataLoss", "FusedAdam", "STCGANEngine", "TrainConfig",
           "infer", "infer_u8", "InferencePipeline", "StcganError", "StcganLibraryError", "Generators", "Discriminators"]

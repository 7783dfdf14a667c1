"""unet_bssfp_b200 -- B200-native (sm_100a) hot path of SomeUserName1/UNet-bSSFP's MRI-conditioned GAN.

Public surface mirrors the reference's seam (ref:src/model.py:15-92): ``Generator``,
``Discriminator``, ``DownSampleConv``, the L1 / BCE-with-logits losses and the evaluation's
relative-error reduction, all running on hand-written CUDA through ``libubssfp.so``.
"""
__version__ = "0.1.0"

from .modules import (BasicUNet, BCEWithLogitsLoss, Discriminator, DownSampleConv, Generator,  # noqa: E402,F401
                      L1Loss, invalidate_packed_weights, set_precision)
from . import ops  # noqa: E402,F401
from . import inference  # noqa: E402,F401
from . import nifti  # noqa: E402,F401
from . import hostmem  # noqa: E402,F401
from .optim import FusedAdamW  # noqa: E402,F401

"""hierarchical_vision_b200 -- B200 (sm_100a) implementation of the SwinV2 windowed-attention hot path
of samuelstevens/hierarchical-vision, behind the reference's own ``swinv2.py`` module API.

    from hierarchical_vision_b200 import swinv2            # drop-in for the reference's swinv2 module
    from hierarchical_vision_b200 import functional        # autograd ops over the C ABI (include/hv_swin.h)
"""
from . import _lib, functional, swinv2  # noqa: F401
from .swinv2 import (BasicLayer, Checkpoint, Mlp, MultitaskHead, PatchEmbed, PatchMerging,  # noqa: F401
                     SwinTransformerBlock, SwinTransformerV2, WindowAttention, swinv2_base, swinv2_tiny,
                     window_partition, window_reverse)

__version__ = "0.1.0"

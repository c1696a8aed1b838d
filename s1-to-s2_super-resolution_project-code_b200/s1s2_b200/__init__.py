"""s1s2_b200 -- B200-native (sm_100a) S1->S2 diffusion sampling path.

Python mirror of the reference's call pattern over the C ABI of libs1s2_b200.so (include/s1s2_b200.h):
``UNetSmallB200`` is the drop-in for ``UNetSmall``; ``samplers`` holds the reference's sampler functions;
``patch`` holds Patch.py's tiling plus the stitch; ``scene`` shards whole scenes across GPUs.
There is no CPU / PyTorch fallback anywhere in this package.
"""
from . import _lib, schedule, metrics, patch, patchio, samplers, scene  # noqa: F401
from ._lib import S1S2Error  # noqa: F401
from .model import UNetSmallB200, synthetic_checkpoint  # noqa: F401
from .schedule import cosine_beta_schedule, linear_beta_schedule, make_schedule  # noqa: F401

__all__ = ["UNetSmallB200", "synthetic_checkpoint", "S1S2Error", "schedule", "metrics", "patch", "samplers", "scene", "cosine_beta_schedule",
           "linear_beta_schedule", "make_schedule"]

import numpy as np   # noqa: F401  (the reference entry point picks np/torch up through its star imports)
import torch         # noqa: F401

from .temporalunet import TemporalUNet, unet_key_table   # noqa: F401
from .diffusion import Diffusion                          # noqa: F401

__all__ = ["np", "torch", "TemporalUNet", "Diffusion", "unet_key_table"]

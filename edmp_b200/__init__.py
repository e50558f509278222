"""edmp_b200 -- B200-native guided-diffusion trajectory sampler (drop-in for the hot path of
vishal-2000/EDMP's infer_serial.py).  Python here is host glue only: weight/scene loading and
the reference-shaped classes; all arithmetic of the hot path runs in libedmp_b200.so
(hand-written sm_100a CUDA behind the C ABI declared in include/edmp_b200.h).
"""
from .diffusion import Diffusion, TemporalUNet          # noqa: F401
from .lib import IntersectionVolumeGuide                 # noqa: F401
from .guide_cfg import Guide, YamlConfig, build_guide_cfgs, load_guide_hparams  # noqa: F401
from . import scene, synthetic                           # noqa: F401

__version__ = "0.1.0"

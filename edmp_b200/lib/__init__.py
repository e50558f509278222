import numpy as np   # noqa: F401
import torch         # noqa: F401

from .guide import IntersectionVolumeGuide                # noqa: F401
from .environment import RobotEnvironment                 # noqa: F401
from .metrics import MetricsCalculator                    # noqa: F401
from .sdf_guide import SphereSDFGuide                     # noqa: F401

__all__ = ["np", "torch", "IntersectionVolumeGuide", "RobotEnvironment", "MetricsCalculator", "SphereSDFGuide"]

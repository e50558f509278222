"""MetricsCalculator name kept for the reference entry point (infer_serial.py:114); the reference
constructs it but never calls its methods there.  OUT OF SCOPE for this tier."""
import numpy as np


class MetricsCalculator:
    def __init__(self, guide):
        self.guide = guide

    def path_length_joint_space(self, trajectory):
        """sum of joint-space segment lengths of a [7, n] trajectory"""
        return float(np.sum(np.linalg.norm(np.diff(np.asarray(trajectory), axis=1), axis=0)))

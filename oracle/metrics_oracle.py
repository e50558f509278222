"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's trajectory metrics.

Each function cites the reference lines it follows:
  * ee_transforms        lib/guide.py:100-116 (get_end_effector_transform: product of the 10 DH matrices, float32;
                         DH matrix lib/guide.py:45-72, static table :29-38)
  * path_lengths         lib/metrics.py:33-45
  * speed_profiles       lib/metrics.py:11-28
  * sparc                lib/metrics.py:47-130 (modified spectral arc length)
Pinned (tests/test_oracle_golden.py) against the UNMODIFIED reference MetricsCalculator + IntersectionVolumeGuide run
in the build container (oracle/make_golden_metrics.py -> tests/golden/metrics.npz).  The product never imports this.
"""
import math

import numpy as np
import torch

from .guide_oracle import _dh_matrix_torch

# (a, d, alpha, theta offset), lib/guide.py:29-38
STATIC_DH = [(0.0, 0.333, 0.0, 0.0), (0.0, 0.0, -math.pi / 2, 0.0), (0.0, 0.316, math.pi / 2, 0.0),
             (0.0825, 0.0, math.pi / 2, 0.0), (-0.0825, 0.384, -math.pi / 2, 0.0), (0.0, 0.0, math.pi / 2, 0.0),
             (0.088, 0.0, math.pi / 2, 0.0), (0.0, 0.107, 0.0, 0.0), (0.0, 0.0, 0.0, -math.pi / 4),
             (0.0, 0.1034, 0.0, 0.0)]


def ee_transforms(joints):
    """joints [B, n, 7] -> float32 [B, n, 4, 4]: T = DH_0(q_0) ... DH_6(q_6) DH_7 DH_8 DH_9 (lib/guide.py:100-116)."""
    q = torch.as_tensor(np.asarray(joints), dtype=torch.float32)
    T = torch.eye(4, dtype=torch.float32).expand(q.shape[0], q.shape[1], 4, 4).clone()
    for i, (a, d, alpha, theta) in enumerate(STATIC_DH):
        ang = q[:, :, i] if i < 7 else torch.full(q.shape[:2], theta, dtype=torch.float32)
        c = lambda v: torch.full(q.shape[:2], v, dtype=torch.float32)   # noqa: E731
        T = torch.matmul(T, _dh_matrix_torch(c(a), c(d), c(alpha), ang))
    return T.numpy()


def _ee_positions(joints_7n):
    """(7, n) float64 -> (n, 3) float32 end-effector positions (lib/metrics.py:17-21, :35-37)."""
    q = np.asarray(joints_7n, dtype=np.float64).T[None].astype(np.float32)
    return ee_transforms(q)[0, :, :3, 3]


def path_lengths(joints_7n):
    """(7, n) -> (joint path length, end-effector path length) (lib/metrics.py:33-45)."""
    pts = _ee_positions(joints_7n)
    wp = np.asarray(joints_7n, dtype=np.float64).T
    ee = np.sum(np.linalg.norm(np.diff(pts, 1, axis=0), axis=1))
    jl = np.sum(np.linalg.norm(np.diff(wp, 1, axis=0), axis=1))
    return float(jl), float(ee)


def speed_profiles(joints_7n, dt):
    """(7, n), dt -> (joint speed [n-1] float64, end-effector speed [n-1] float32) (lib/metrics.py:24-28)."""
    pts = _ee_positions(joints_7n)
    wp = np.asarray(joints_7n, dtype=np.float64).T
    return (np.linalg.norm(np.diff(wp, n=1, axis=0) / dt, axis=1),
            np.linalg.norm(np.diff(pts, n=1, axis=0) / dt, axis=1))


def sparc(movement, fs, padlevel=4, fc=10.0, amp_th=0.05):
    """-> (sal, f, Mf, first, last): spectral arc length of the normalised magnitude spectrum of the zero-padded
    profile, between the first and last bin at or above amp_th among the bins with f <= fc (lib/metrics.py:86-130).
    All-zero profile -> (0, None, None, -1, -1) (:86-88)."""
    movement = np.asarray(movement)
    if np.allclose(movement, 0):
        return 0.0, None, None, -1, -1
    nfft = int(2 ** (np.ceil(np.log2(len(movement))) + padlevel))          # :90
    f = np.arange(0, fs, fs / nfft)                                         # :93
    Mf = np.abs(np.fft.fft(movement, nfft))                                 # :95
    Mf = Mf / Mf.max()                                                      # :96
    low = np.flatnonzero(f <= fc)                                           # :103-105
    f_lp, Mf_lp = f[low], Mf[low]
    above = np.flatnonzero(Mf_lp >= amp_th)                                 # :110
    first, last = int(above[0]), int(above[-1])
    f_sel, Mf_sel = f_lp[first:last + 1], Mf_lp[first:last + 1]             # :111-113
    arc = np.sqrt((np.diff(f_sel) / (f_sel[-1] - f_sel[0])) ** 2 + np.diff(Mf_sel) ** 2) if last > first else np.zeros(0)
    return float(-np.sum(arc)), f, Mf, int(low[first]), int(low[last])     # :116-121


def trajectory_metrics(joints_7n, dt, padlevel=4, fc=10.0, amp_th=0.05):
    """(7, n), dt -> [joint path length, ee path length, joint SPARC, ee SPARC] (lib/metrics.py:11-45)."""
    jl, el = path_lengths(joints_7n)
    vj, ve = speed_profiles(joints_7n, dt)
    return np.array([jl, el, sparc(vj, 1. / dt, padlevel, fc, amp_th)[0], sparc(ve, 1. / dt, padlevel, fc, amp_th)[0]])

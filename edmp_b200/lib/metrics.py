"""MetricsCalculator -- reference-shaped boundary class over the CUDA metrics kernels.

Same constructor and method signatures as reference lib/metrics.py:4-130: ``smoothness_metric(joints, dt)``,
``path_length_metric(joints)``, ``sparc(movement, fs, padlevel=4, fc=10.0, amp_th=0.05)`` with the same return
shapes (``sparc`` returns ``(sal, (f, Mf), (f_sel, Mf_sel))`` and ``smoothness_metric`` returns two such tuples, as the
reference does).  The reference evaluates one [7, 50] trajectory per call on the host; ``ensemble_metrics`` is the
batched form: every row of a sampled ensemble in one launch of libedmp_b200.so (edmp_trajectory_metrics).
There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


class MetricsCalculator:
    def __init__(self, guide):
        self.guide = guide
        self.device = self.guide.device

    # ---- batched engine calls --------------------------------------------------------------------
    def ensemble_metrics(self, trajectories, dt, padlevel=4, fc=10.0, amp_th=0.05, return_spectra=False):
        """trajectories [B, 7, n] (numpy or tensor) -> dict of float64 [B] arrays: joint_path_length,
        end_eff_path_length, joint_smoothness, end_eff_smoothness (SPARC of the speed profiles at fs = 1 / dt).
        With return_spectra also ``spectra`` [B, 2, nfft] (normalised magnitude, joint / end effector) and
        ``selected`` [B, 2, 2] (first, last bin of the arc; -1 = empty)."""
        dev = _lib.require_cuda(self.device)
        traj = torch.as_tensor(np.asarray(trajectories, dtype=np.float64) if not torch.is_tensor(trajectories)
                               else trajectories).to(dev, torch.float64).contiguous()
        if traj.dim() != 3 or traj.shape[1] != 7:
            raise ValueError("trajectories must be [B, 7, n]")
        rows, n = int(traj.shape[0]), int(traj.shape[2])
        lib = _lib.load()
        nfft = lib.edmp_metrics_nfft(n - 1, int(padlevel))
        with torch.cuda.device(dev):
            out = torch.empty(rows, 4, device=dev, dtype=torch.float64)
            spec = torch.empty(rows, 2, max(nfft, 1), device=dev, dtype=torch.float64) if return_spectra else None
            sel = torch.empty(rows, 2, 2, device=dev, dtype=torch.int32) if return_spectra else None
            _lib.check(lib.edmp_trajectory_metrics(ctypes.c_void_p(traj.data_ptr()), rows, n, float(dt), int(padlevel),
                                                   float(fc), float(amp_th), ctypes.c_void_p(out.data_ptr()),
                                                   ctypes.c_void_p(spec.data_ptr()) if return_spectra else None,
                                                   ctypes.c_void_p(sel.data_ptr()) if return_spectra else None,
                                                   _lib.stream_ptr()), "edmp_trajectory_metrics")
        o = out.cpu().numpy()
        res = {"joint_path_length": o[:, 0], "end_eff_path_length": o[:, 1], "joint_smoothness": o[:, 2],
               "end_eff_smoothness": o[:, 3]}
        if return_spectra:
            res["spectra"] = spec.cpu().numpy()
            res["selected"] = sel.cpu().numpy()
            res["nfft"] = nfft
        return res

    @staticmethod
    def _sparc_tuple(sal, Mf, sel, fs, nfft):
        """(sal, (f, Mf), (f_sel, Mf_sel)) of reference lib/metrics.py:130; (0, None, None) for an all-zero profile
        (:86-88)."""
        if sel[1] < 0 and not np.any(Mf):
            return 0, None, None
        f = np.arange(0, fs, fs / nfft)
        lo, hi = int(sel[0]), int(sel[1])
        return sal, (f, Mf), (f[lo:hi + 1], Mf[lo:hi + 1])

    # ---- reference API ---------------------------------------------------------------------------
    def smoothness_metric(self, joints, dt):
        """joints: (7, n) array, dt: time between waypoints -> (joint SPARC tuple, end-effector SPARC tuple)
        (reference lib/metrics.py:11-31)."""
        r = self.ensemble_metrics(np.asarray(joints)[None], dt, return_spectra=True)
        fs = 1. / dt
        out = []
        for which, key in enumerate(("joint_smoothness", "end_eff_smoothness")):
            out.append(self._sparc_tuple(float(r[key][0]), r["spectra"][0, which], r["selected"][0, which], fs, r["nfft"]))
        return out[0], out[1]

    def path_length_metric(self, joints):
        """joints: (7, n) array -> (joint path length, end-effector path length) (reference lib/metrics.py:33-45)."""
        r = self.ensemble_metrics(np.asarray(joints)[None], 1.0)
        return float(r["joint_path_length"][0]), float(r["end_eff_path_length"][0])

    def sparc(self, movement, fs, padlevel=4, fc=10.0, amp_th=0.05):
        """Spectral arc length of a 1-D speed profile (reference lib/metrics.py:47-130), evaluated on the device."""
        dev = _lib.require_cuda(self.device)
        mv = torch.as_tensor(np.asarray(movement, dtype=np.float64)).reshape(1, -1).to(dev).contiguous()
        m = int(mv.shape[1])
        lib = _lib.load()
        nfft = lib.edmp_metrics_nfft(m, int(padlevel))
        with torch.cuda.device(dev):
            sal = torch.empty(1, device=dev, dtype=torch.float64)
            spec = torch.empty(1, max(nfft, 1), device=dev, dtype=torch.float64)
            sel = torch.empty(1, 2, device=dev, dtype=torch.int32)
            _lib.check(lib.edmp_sparc(ctypes.c_void_p(mv.data_ptr()), 1, m, float(fs), int(padlevel), float(fc),
                                      float(amp_th), ctypes.c_void_p(sal.data_ptr()), ctypes.c_void_p(spec.data_ptr()),
                                      ctypes.c_void_p(sel.data_ptr()), _lib.stream_ptr()), "edmp_sparc")
        return self._sparc_tuple(float(sal.cpu()[0]), spec.cpu().numpy()[0], sel.cpu().numpy()[0], fs, nfft)

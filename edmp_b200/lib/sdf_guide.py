"""Sphere / signed-distance guide family (host side; SURVEY.md section 8 a-S, BASELINE.json configs[4]).

Analytic Franka FK -> the 59 robofin collision spheres -> signed distance to box / cylinder primitives (mpinets
geometry semantics) or nearest-point distance to a point cloud -> hinge cost, analytic gradient, per-waypoint
clearance.  All arithmetic runs in libedmp_b200.so (csrc/sdf_guide.cu); this class only owns handles and tensors.
It is a second guide family next to IntersectionVolumeGuide (the one the reference's infer_serial.py uses) and is
not wired into the sampler loop yet.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


class SphereSDFGuide:
    """boxes: [nb,10] = (xyz, quaternion xyzw, dims) like the reference's obstacle_config rows;
    cylinders: [nc,9] = (xyz, quaternion xyzw, radius, height); at most 64 primitives in total."""

    def __init__(self, boxes=None, cylinders=None, device="cuda:0", margin=0.03):
        self.device = device
        self.margin = float(margin)          # mpinets/loss.py:88-94 hinge margin
        self.boxes = np.ascontiguousarray(np.zeros((0, 10)) if boxes is None else boxes, dtype=np.float64).reshape(-1, 10)
        self.cylinders = np.ascontiguousarray(np.zeros((0, 9)) if cylinders is None else cylinders,
                                              dtype=np.float64).reshape(-1, 9)
        dev = _lib.require_cuda(device)
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.load().edmp_sdf_scene_create(
                self.boxes.ctypes.data_as(ctypes.c_void_p), self.boxes.shape[0],
                self.cylinders.ctypes.data_as(ctypes.c_void_p), self.cylinders.shape[0], ctypes.byref(handle)),
                "edmp_sdf_scene_create")
        self._scene = handle

    def __del__(self):
        try:
            if getattr(self, "_scene", None) is not None:
                _lib.load().edmp_sdf_scene_destroy(self._scene)
        except Exception:
            pass

    def evaluate(self, q, want_grad=True, want_clearance=True):
        """q: [B,7,n] joint trajectories -> (cost [B], grad [B,7,n] | None, clearance [B,n] | None), CUDA float32."""
        dev = _lib.require_cuda(self.device)
        q = torch.as_tensor(q).to(dev, torch.float32).contiguous()
        if q.dim() != 3 or q.shape[1] != 7 or not 1 <= q.shape[2] <= 64:
            raise ValueError("expected q of shape [B, 7, n <= 64], got %s" % (tuple(q.shape),))
        B, _, n = q.shape
        cost = torch.empty(B, device=dev, dtype=torch.float32)
        grad = torch.empty_like(q) if want_grad else None
        clr = torch.empty(B, n, device=dev, dtype=torch.float32) if want_clearance else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().edmp_sdf_guide(
                self._scene, ctypes.c_void_p(q.data_ptr()), n, B, ctypes.c_float(self.margin),
                ctypes.c_void_p(cost.data_ptr()), ctypes.c_void_p(grad.data_ptr()) if want_grad else None,
                ctypes.c_void_p(clr.data_ptr()) if want_clearance else None, _lib.stream_ptr()), "edmp_sdf_guide")
        return cost, grad, clr

    def cost(self, q):
        return self.evaluate(q, False, False)[0]

    def get_gradient(self, q):
        return self.evaluate(q, True, False)[1]

    def clearance(self, q):
        return self.evaluate(q, False, True)[2]

    def has_collision(self, q):
        """[B,7,n] -> bool [B] (CUDA): the rollout collision predicate of the reference's validation step
        (mpinets/model.py:296-312): some collision sphere at some waypoint has scene sdf <= its radius, i.e. the
        trajectory's minimum clearance is <= 0.  One launch for the whole ensemble."""
        return (self.clearance(q) <= 0.0).any(dim=1)

    def cloud_clearance(self, points, q):
        """points: [P,3] scene point cloud -> [B,n] nearest (sphere surface, point) distance per waypoint."""
        dev = _lib.require_cuda(self.device)
        q = torch.as_tensor(q).to(dev, torch.float32).contiguous()
        pts = torch.as_tensor(points).to(dev, torch.float32)
        if pts.dim() != 2 or pts.shape[1] != 3 or pts.shape[0] == 0:
            raise ValueError("expected a non-empty [P, 3] point cloud")
        p4 = torch.zeros(pts.shape[0], 4, device=dev, dtype=torch.float32)
        p4[:, :3] = pts
        B, _, n = q.shape
        if n > 50:
            raise ValueError("the point-cloud kernel takes at most 50 waypoints")
        out = torch.empty(B, n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().edmp_sdf_cloud_clearance(
                ctypes.c_void_p(q.data_ptr()), n, B, ctypes.c_void_p(p4.data_ptr()), p4.shape[0],
                ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr()), "edmp_sdf_cloud_clearance")
        return out

"""IntersectionVolumeGuide -- reference-shaped boundary class over the CUDA guide kernels.

Same constructor and method signatures as reference lib/guide.py:11-677 for the methods
infer_serial.py, the sampler and lib/metrics.py use: ``cost`` (:354), ``swept_volume_cost`` (:473),
``get_gradient`` (:597), ``choose_best_trajectory`` (:637), ``rearrange_joints``.  The gradient is
the closed form of the reference's autograd result, evaluated in libedmp_b200.so.
"""
import ctypes
import os
import re
import warnings

import numpy as np
import torch

from .. import _lib

LINK_NAMES = ["link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand", "finger"]


def link_dimensions_from_meshes(folder):
    """AABB extents of the 9 collision meshes the way the reference measures them
    (lib/guide.py:243-281): max-min over OBJ 'v' records, finger y extent x4."""
    dims = []
    for i, name in enumerate(LINK_NAMES):
        verts = []
        with open(os.path.join(folder, name + ".obj"), "r") as f:
            for line in f:
                line = line.strip()
                if line.startswith("v "):
                    verts.append([float(c) for c in re.split(r"\s+", line)[1:4]])
        v = np.array(verts)
        d = v.max(axis=0) - v.min(axis=0)
        if i == len(LINK_NAMES) - 1:
            d[1] *= 4
        dims.append(d)
    return np.array(dims)


#: where the link-box extents of the last guide came from: "pybullet_data" (what the reference measures,
#: lib/guide.py:245) or "builtin" (the library's table: robofin's hd collision meshes, SURVEY.md 8c -- the same
#: geometry at another tessellation, extents differ by up to a few mm from pybullet_data's)
LINK_DIMENSIONS_SOURCE = None
_warned_builtin = False


def _default_link_dimensions():
    global LINK_DIMENSIONS_SOURCE, _warned_builtin
    try:
        import pybullet_data
    except ImportError:
        pybullet_data = None
    if pybullet_data is not None:
        folder = os.path.join(pybullet_data.getDataPath(), "franka_panda", "meshes", "collision")
        if os.path.isdir(folder):
            LINK_DIMENSIONS_SOURCE = "pybullet_data"
            return link_dimensions_from_meshes(folder)   # a broken mesh folder raises: do not hide it
    LINK_DIMENSIONS_SOURCE = "builtin"
    if not _warned_builtin:
        _warned_builtin = True
        warnings.warn("edmp_b200: pybullet_data is not installed; link-box extents come from the built-in table "
                      "(robofin hd collision meshes), which differs from a stock reference install by up to a few mm. "
                      "Pass link_dimensions= to IntersectionVolumeGuide to pin them.", RuntimeWarning, stacklevel=3)
    return None


class IntersectionVolumeGuide:

    def __init__(self, obstacle_config, device, guide_cfgs, batch_size, link_dimensions=None):
        self.device = device
        self.guide_cfgs = guide_cfgs
        self.obstacle_config = np.array(obstacle_config, dtype=np.float64)
        self.batch_size = batch_size
        if self.obstacle_config.ndim != 2 or self.obstacle_config.shape[1] != 10:
            raise ValueError("obstacle_config must be [n, 10] = (xyz, quat xyzw, dims)")
        self.link_dimensions = link_dimensions if link_dimensions is not None else _default_link_dimensions()
        self._scene = None
        self._tables_key = None
        #: rows per ensemble when one batch holds several (None = the whole batch is one ensemble)
        self.ensemble_rows = None

    def rearrange_joints(self, x):
        # 'batch channels traj_len -> batch traj_len channels' (lib/guide.py:43)
        return x.permute(0, 2, 1)

    def get_end_effector_transform(self, joints):
        """[b, n, 7] joint tensor -> float32 [b, n, 4, 4] end-effector transforms on the guide's device: the product
        of the 10 DH matrices (reference lib/guide.py:100-116)."""
        dev = _lib.require_cuda(self.device)
        q = torch.as_tensor(joints).to(dev, torch.float32).contiguous()
        if q.dim() != 3 or q.shape[2] != 7:
            raise ValueError("joints must be [batch, traj_len, 7]")
        b, n = int(q.shape[0]), int(q.shape[1])
        with torch.cuda.device(dev):
            T = torch.empty(b, n, 4, 4, device=dev, dtype=torch.float32)
            _lib.check(_lib.load().edmp_ee_transform(ctypes.c_void_p(q.data_ptr()), b, n,
                                                     ctypes.c_void_p(T.data_ptr()), _lib.stream_ptr()),
                       "edmp_ee_transform")
        return T

    # ---- engine --------------------------------------------------------------------------------------
    def _scene_handle(self):
        if self._scene is None:
            dev = _lib.require_cuda(self.device)
            lib = _lib.load()
            cfg, cfg_ptr = _lib.host_f64(self.obstacle_config)
            ld_ptr = None
            if self.link_dimensions is not None:
                ld, ld_ptr = _lib.host_f64(self.link_dimensions)
            handle = ctypes.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(lib.edmp_scene_create(cfg_ptr, cfg.shape[0], ld_ptr, ctypes.byref(handle)),
                           "edmp_scene_create")
            self._scene = handle
        return self._scene

    def scene_handle(self, rows=None, guidance_schedule=None, ensemble_rows=None):
        """edmp_scene* with the per-row guide tables (infer_serial.py:59-91) uploaded for ``rows``."""
        scene = self._scene_handle()
        rows = int(rows if rows is not None else self.batch_size)
        sched = guidance_schedule if guidance_schedule is not None else self.guide_cfgs["guidance_schedule"]
        if ensemble_rows is None:
            ensemble_rows = self.ensemble_rows
        ens = int(ensemble_rows if ensemble_rows is not None else rows)
        g = self.guide_cfgs
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in
                (g["clearance"], g["expansion"], sched, g["guidance_method"], g["grad_norm"])]
        # keyed on the identity AND a content fingerprint of all five tables (the arrays are kept referenced, so an id()
        # cannot be recycled by a temporary; the two sums catch in-place edits of guide_cfgs between calls at ~1 ms per
        # 2M-entry table -- a cryptographic hash of the 50 MB of an 8190-row ensemble would cost 50 ms per call)
        srcs = (g["clearance"], g["expansion"], sched, g["guidance_method"], g["grad_norm"])
        finger = tuple((float(a.sum()), float(a.reshape(-1)[::13].sum())) for a in arrs)
        key = (rows, ens, tuple(id(o) for o in srcs), finger)
        self._tables_refs = srcs
        if self._tables_key != key:
            for a, width in zip(arrs, (255, 255, 255, None, None)):
                if a.shape[0] != rows or (width and a.shape[1] != width):
                    raise ValueError("guide table shape %s does not match rows=%d" % (a.shape, rows))
            with torch.cuda.device(torch.device(self.device)):
                _lib.check(_lib.load().edmp_scene_set_guide_tables(
                    scene, *[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], rows, ens),
                    "edmp_scene_set_guide_tables")
            self._tables_key = key
        return scene

    def __del__(self):
        try:
            if self._scene is not None:
                _lib.load().edmp_scene_destroy(self._scene)
        except Exception:
            pass

    def _volumes(self, joint_input, start, goal, t, mode, batch_size):
        dev = _lib.require_cuda(self.device)
        q = torch.as_tensor(joint_input).to(dev, torch.float32).contiguous()
        rows, _, n = q.shape
        with torch.cuda.device(dev):
            scene = self.scene_handle(rows=rows) if t != 0 else self._scene_handle()
            no = self.obstacle_config.shape[0]
            out = torch.empty(rows, n + 1 if mode else n, 9 * no, device=dev, dtype=torch.float32)
            s_ptr = g_ptr = None
            if mode:
                s_arr, s_ptr = _lib.host_f64(torch.as_tensor(start).detach().cpu().numpy())
                g_arr, g_ptr = _lib.host_f64(torch.as_tensor(goal).detach().cpu().numpy())
                if g_arr.size != 7:
                    raise NotImplementedError("per-row goals are not used by infer_serial.py")
            _lib.check(_lib.load().edmp_guide_volumes(scene, ctypes.c_void_p(q.data_ptr()), s_ptr, g_ptr, int(t),
                                                      mode, rows, n, ctypes.c_void_p(out.data_ptr()),
                                                      _lib.stream_ptr()), "edmp_guide_volumes")
        return out

    # ---- reference API -------------------------------------------------------------------------------
    def cost(self, joint_input, t, batch_size=None):
        """[b, 7, n] -> intersection volumes [b, n, 9*n_obs] (index link*n_obs + obstacle)."""
        return self._volumes(joint_input, None, None, t, 0, batch_size)

    def swept_volume_cost(self, joint_input, start, goal, t, batch_size=None):
        """[b, 7, n] -> swept volumes [b, n+1, 9*n_obs] over [start | joints | goal]."""
        return self._volumes(joint_input, start, goal, t, 1, batch_size)

    def get_gradient(self, joint_input, start, goal, t, return_raw=False):
        """np [B, 7, 48] (already clipped) -> np.float64 [B, 7, 48]."""
        dev = _lib.require_cuda(self.device)
        q = torch.as_tensor(np.asarray(joint_input, dtype=np.float64)).to(dev).contiguous()
        rows = q.shape[0]
        if tuple(q.shape[1:]) != (7, 48):
            raise ValueError("joint_input must be [B, 7, 48]")
        with torch.cuda.device(dev):
            scene = self.scene_handle(rows=rows)
            grad = torch.empty_like(q)
            raw = torch.empty(q.shape, device=dev, dtype=torch.float32) if return_raw else None
            s_arr, s_ptr = _lib.host_f64(start)
            g_arr, g_ptr = _lib.host_f64(goal)
            _lib.check(_lib.load().edmp_guide_gradient(scene, ctypes.c_void_p(q.data_ptr()), s_ptr, g_ptr, int(t),
                                                       rows, ctypes.c_void_p(grad.data_ptr()),
                                                       ctypes.c_void_p(raw.data_ptr()) if return_raw else None,
                                                       _lib.stream_ptr()), "edmp_guide_gradient")
        if return_raw:
            return grad.cpu().numpy(), raw.cpu().numpy()
        return grad.cpu().numpy()

    def final_costs(self, start, goal, trajectories):
        """per-row sum of swept volumes at t=0, np.float32 [B]"""
        dev = _lib.require_cuda(self.device)
        traj = torch.as_tensor(np.asarray(trajectories, dtype=np.float64)).to(dev).contiguous()
        with torch.cuda.device(dev):
            cost = torch.empty(traj.shape[0], device=dev, dtype=torch.float32)
            s_arr, s_ptr = _lib.host_f64(start)
            g_arr, g_ptr = _lib.host_f64(goal)
            _lib.check(_lib.load().edmp_guide_final_cost(self._scene_handle(), ctypes.c_void_p(traj.data_ptr()),
                                                         s_ptr, g_ptr, traj.shape[0],
                                                         ctypes.c_void_p(cost.data_ptr()), _lib.stream_ptr()),
                       "edmp_guide_final_cost")
        return cost.cpu().numpy()

    def choose_best_trajectory(self, start, goal, trajectories):
        costs = self.final_costs(start, goal, trajectories)
        return trajectories[int(np.argmin(costs))]

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the sphere / signed-distance guide family (SURVEY.md section 8 a-S).

Each function cites the reference lines it follows (code the reference vendors but its infer_serial.py never calls):
  * SPHERES            robofin/robofin/robots.py:58-174
  * franka_fk          robofin/robofin/urdf/franka_panda/panda.urdf:47-235,:310-324 evaluated like
                       robofin/robofin/torch_urdf.py:466-522 (origin xyz/rpy, then the joint rotation about z)
  * sphere_centres     robofin/robofin/pointcloud/torch.py:340-365 (compute_spheres)
  * box_sdf            mpinets/geometry.py:238-288   (quaternion convention: obstacle_config rows are xyzw,
                       datasets/load_test_dataset.py:189; mpinets stores wxyz, geometry.py:60-63)
  * cylinder_sdf       mpinets/geometry.py:456-505
  * cost               builder-defined hinge on (radius + margin - sdf), margin 0.03 as in mpinets/loss.py:88-94;
                       collision predicate sdf <= radius as in mpinets/model.py:301-312
Pinned (tests/test_oracle_golden.py): box_sdf / cylinder_sdf against the reference's own TorchCuboids / TorchCylinders
(imported from /root/reference under a geometrout stub, fixtures committed in tests/golden/sdf.npz); the FK against
the reference-pinned DH chain of lib/guide.py (same joint origins and axes).  The guided-sampler use of this family
and the point-cloud variant have no reference implementation: "parity unpinned" for those.
"""
import math

import numpy as np
import torch

SPHERES = [
    (0.08, {"panda_link0": [[0.0, 0.0, 0.05]]}),
    (0.06, {"panda_link1": [[0.0, -0.08, 0.0], [0.0, -0.03, 0.0], [0.0, 0.0, -0.12], [0.0, 0.0, -0.17]],
            "panda_link2": [[0.0, 0.0, 0.03], [0.0, 0.0, 0.08], [0.0, -0.12, 0.0], [0.0, -0.17, 0.0]],
            "panda_link3": [[0.0, 0.0, -0.1]],
            "panda_link4": [[-0.08, 0.095, 0.0]],
            "panda_link5": [[0.0, 0.055, 0.0], [0.0, 0.075, 0.0], [0.0, 0.0, -0.22]]}),
    (0.05, {"panda_link3": [[0.0, 0.0, -0.06]], "panda_link5": [[0.0, 0.05, -0.18]],
            "panda_link6": [[0.0, 0.0, 0.0], [0.08, -0.01, 0.0]], "panda_link7": [[0.0, 0.0, 0.07]]}),
    (0.055, {"panda_link3": [[0.08, 0.06, 0.0], [0.08, 0.02, 0.0]],
             "panda_link4": [[0.0, 0.0, 0.02], [0.0, 0.0, 0.06], [-0.08, 0.06, 0.0]]}),
    (0.025, {"panda_link5": [[0.01, 0.08, -0.14], [0.01, 0.085, -0.11], [0.01, 0.09, -0.08], [0.01, 0.095, -0.05],
                             [-0.01, 0.08, -0.14], [-0.01, 0.085, -0.11], [-0.01, 0.09, -0.08], [-0.01, 0.095, -0.05]],
             "panda_link7": [[0.02, 0.04, 0.08], [0.04, 0.02, 0.08]]}),
    (0.052, {"panda_link6": [[0.08, 0.035, 0.0]]}),
    (0.02, {"panda_link7": [[0.04, 0.06, 0.085], [0.06, 0.04, 0.085]]}),
    (0.028, {"panda_hand": [[0.0, y, 0.01] for y in (-0.075, -0.045, -0.015, 0.015, 0.045, 0.075)]}),
    (0.026, {"panda_hand": [[0.0, y, 0.03] for y in (-0.075, -0.045, -0.015, 0.015, 0.045, 0.075)]}),
    (0.024, {"panda_hand": [[0.0, y, 0.05] for y in (-0.075, -0.045, -0.015, 0.015, 0.045, 0.075)]}),
    (0.012, {"panda_leftfinger": [[0, 0.015, 0.022], [0, 0.008, 0.044]],
             "panda_rightfinger": [[0, -0.015, 0.022], [0, -0.008, 0.044]]}),
]

# (parent, xyz, roll) of the seven revolute joints, panda.urdf:47-217 (pitch = yaw = 0 everywhere)
_JOINTS = [((0.0, 0.0, 0.333), 0.0), ((0.0, 0.0, 0.0), -math.pi / 2), ((0.0, -0.316, 0.0), math.pi / 2),
           ((0.0825, 0.0, 0.0), math.pi / 2), ((-0.0825, 0.384, 0.0), -math.pi / 2), ((0.0, 0.0, 0.0), math.pi / 2),
           ((0.088, 0.0, 0.0), math.pi / 2)]


def _homog(R, t, like):
    T = torch.zeros(*like.shape[:-1], 4, 4, dtype=like.dtype)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


def _rx(a, dtype):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[1.0, 0, 0], [0, c, -s], [0, s, c]], dtype=dtype)


def _rz(q):
    c, s = torch.cos(q), torch.sin(q)
    z, o = torch.zeros_like(q), torch.ones_like(q)
    return torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)


def franka_fk(q):
    """q [..., 7] -> dict link name -> [..., 4, 4] world frames (link0 = identity)."""
    q = torch.as_tensor(q)
    dtype = q.dtype
    eye = torch.eye(4, dtype=dtype).expand(*q.shape[:-1], 4, 4)
    frames = {"panda_link0": eye}
    T = eye
    for i, (xyz, roll) in enumerate(_JOINTS):
        R = _rx(roll, dtype) @ _rz(q[..., i])
        T = T @ _homog(R, torch.tensor(xyz, dtype=dtype), q)
        frames["panda_link%d" % (i + 1)] = T
    zero = torch.zeros_like(q[..., 0])
    link8 = T @ _homog(torch.eye(3, dtype=dtype), torch.tensor([0.0, 0.0, 0.107], dtype=dtype), q)   # panda.urdf:225-230
    hand = link8 @ _homog(_rz(zero - math.pi / 4), torch.zeros(3, dtype=dtype), q)                        # :231-235
    frames["panda_hand"] = hand
    # prismatic fingers at 0.025 (pointcloud/torch.py:343-350), axes +y / -y (panda.urdf:310-324)
    frames["panda_leftfinger"] = hand @ _homog(torch.eye(3, dtype=dtype), torch.tensor([0.0, 0.025, 0.0584], dtype=dtype), q)
    frames["panda_rightfinger"] = hand @ _homog(torch.eye(3, dtype=dtype), torch.tensor([0.0, -0.025, 0.0584], dtype=dtype), q)
    return frames


def sphere_centres(q):
    """q [..., 7] -> (centres [..., 59, 3], radii [59]) in the SPHERES list order."""
    fk = franka_fk(q)
    cs, rs = [], []
    for radius, per_link in SPHERES:
        for link, pts in per_link.items():
            p = torch.tensor(pts, dtype=fk[link].dtype)
            T = fk[link]
            cs.append(torch.einsum("...ij,nj->...ni", T[..., :3, :3], p) + T[..., None, :3, 3])
            rs += [radius] * len(pts)
    return torch.cat(cs, dim=-2), torch.tensor(rs, dtype=cs[0].dtype)


def inverse_rotation_like_reference(quat_xyzw):
    """World -> primitive rotation exactly as mpinets/geometry.py:185-214 builds it from the conjugate quaternion.
    NB the reference's matrix has `yz - wx` in BOTH (1,2) and (2,1) (geometry.py:209-210; a proper rotation has
    `yz + wx` at (2,1)), so for quaternions with w*x != 0 it is not orthogonal.  Yaw-only obstacles (x = y = 0), the
    common case in the MpiNets scenes, are unaffected.  Restated as is: parity is with the reference."""
    qv = np.asarray(quat_xyzw, dtype=np.float64)
    qv = qv / np.linalg.norm(qv)
    w, x, y, z = qv[3], -qv[0], -qv[1], -qv[2]
    xx, yy, zz = 2 * x * x, 2 * y * y, 2 * z * z
    wx, wy, wz, xy, xz, yz = 2 * w * x, 2 * w * y, 2 * w * z, 2 * x * y, 2 * x * z, 2 * y * z
    return np.array([[1 - yy - zz, xy - wz, xz + wy],
                     [xy + wz, 1 - xx - zz, yz - wx],
                     [xz - wy, yz - wx, 1 - xx - yy]])


def _to_local(points, centre, quat_xyzw):
    Rinv = torch.tensor(inverse_rotation_like_reference(quat_xyzw), dtype=points.dtype)
    return (points - torch.tensor(np.asarray(centre), dtype=points.dtype)) @ Rinv.T     # Rinv (x - c)


def box_sdf(points, boxes):
    """points [..., 3], boxes [nb,10] (xyz, quat xyzw, dims) -> [..., nb]"""
    out = []
    for b in np.asarray(boxes, dtype=np.float64).reshape(-1, 10):
        p = _to_local(points, b[0:3], b[3:7])
        d = p.abs() - torch.tensor(b[7:10] / 2, dtype=points.dtype)
        outside = torch.linalg.norm(torch.clamp(d, min=0.0), dim=-1)
        inside = torch.clamp(d.max(dim=-1).values, max=0.0)
        out.append(outside + inside)
    return torch.stack(out, dim=-1)


def cylinder_sdf(points, cyls):
    """points [..., 3], cyls [nc,9] (xyz, quat xyzw, radius, height) -> [..., nc]"""
    out = []
    for c in np.asarray(cyls, dtype=np.float64).reshape(-1, 9):
        p = _to_local(points, c[0:3], c[3:7])
        d2 = torch.stack([torch.linalg.norm(p[..., :2], dim=-1) - c[7], p[..., 2].abs() - c[8] / 2], dim=-1)
        outside = torch.linalg.norm(torch.clamp(d2, min=0.0), dim=-1)
        inside = torch.clamp(d2.max(dim=-1).values, max=0.0)
        out.append(outside + inside)
    return torch.stack(out, dim=-1)


def scene_sdf(points, boxes, cyls):
    parts = []
    if boxes is not None and len(boxes):
        parts.append(box_sdf(points, boxes))
    if cyls is not None and len(cyls):
        parts.append(cylinder_sdf(points, cyls))
    return torch.cat(parts, dim=-1).min(dim=-1).values


def evaluate(q, boxes, cyls, margin=0.03, want_grad=True):
    """q [B,7,n] -> cost [B], grad [B,7,n] (autograd), clearance [B,n]; float64."""
    q = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=want_grad)
    c, r = sphere_centres(q.permute(0, 2, 1))                  # [B,n,59,3]
    sdf = scene_sdf(c, boxes, cyls)                            # [B,n,59]
    cost = torch.clamp(r + margin - sdf, min=0.0).sum(dim=(1, 2))
    grad = None
    if want_grad:
        cost.sum().backward()
        grad = q.grad.detach().numpy()
    return cost.detach().numpy(), grad, (sdf - r).min(dim=-1).values.detach().numpy()


def has_collision(q, boxes, cyls):
    """q [B,7,n] -> bool [B]: the rollout collision predicate of the reference's validation step
    (mpinets/model.py:296-312): per sphere radius, some sphere centre at some waypoint has scene sdf <= radius."""
    q = torch.tensor(np.asarray(q, dtype=np.float64))
    c, r = sphere_centres(q.permute(0, 2, 1))                  # [B,n,59,3], [59]
    sdf = scene_sdf(c, boxes, cyls)                            # [B,n,59]
    hit = torch.zeros(q.shape[0], dtype=torch.bool)
    for radius in sorted(set(r.tolist())):
        sel = r == radius
        hit = torch.logical_or(hit, torch.any(sdf[:, :, sel].reshape(q.shape[0], -1) <= radius, dim=-1))
    return hit.numpy()


def cloud_clearance(q, points):
    """q [B,7,n], points [P,3] -> [B,n] = min over (sphere, point) of |centre - p| - radius"""
    q = torch.tensor(np.asarray(q, dtype=np.float64))
    c, r = sphere_centres(q.permute(0, 2, 1))
    p = torch.tensor(np.asarray(points, dtype=np.float64))
    d = torch.cdist(c.reshape(-1, 3), p).min(dim=1).values.reshape(c.shape[:-1])
    return (d - r).min(dim=-1).values.numpy()


# ---- the reference's own SDF (only where /root/reference exists) -------------------------------------------------
def load_reference_geometry():
    """mpinets.geometry imported from the unmodified reference under a geometrout stub."""
    import importlib.util
    import os
    import sys
    import types
    from . import ref_shim
    if "geometrout" not in sys.modules:
        g = types.ModuleType("geometrout")
        prim = types.ModuleType("geometrout.primitive")
        tr = types.ModuleType("geometrout.transform")
        for name in ("Sphere", "Cuboid", "Cylinder"):
            setattr(prim, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
        for name in ("SE3", "SO3"):
            setattr(tr, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
        g.primitive, g.transform = prim, tr
        sys.modules.update({"geometrout": g, "geometrout.primitive": prim, "geometrout.transform": tr})
    path = os.path.join(ref_shim.REF_ROOT, "mpinets", "geometry.py")
    spec = importlib.util.spec_from_file_location("_ref_mpinets_geometry", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

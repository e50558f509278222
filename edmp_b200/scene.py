"""Scene / goal front-end: the per-scene host steps immediately before the hot path.

  * flatten_obstacles  -- the obstacle flattening of the reference's problem loader
                          (datasets/load_test_dataset.py:90-151): cuboids and cylinders of an MpiNets planning problem
                          -> obstacle_config [no,10] = (xyz, quaternion xyzw, dims), cylinders flattened to boxes with
                          dims (r, r, h) (:136-139), quaternions rolled from the problem sets' wxyz to xyzw (:128,:135).
  * select_goal        -- the IK-goal filter of the entry point (infer_serial.py:119-129): among candidate goal
                          configurations keep those whose t=0 intersection volume is within a trust region of the best,
                          take the one nearest to the start in joint space.  The volumes come from the guide kernel
                          (one batched `cost` call); the candidates themselves are an input (the reference samples them
                          with ikfast, load_test_dataset.py:176-186, which is not rebuilt).
Host glue only (numpy); the only arithmetic of note -- the k x 9 x n_obs overlap volumes -- runs in libedmp_b200.so.
"""
import numpy as np

GOAL_TRUST_REGION = 0.0008                 # reference infer_serial.py:125 (hard-coded there too)


def _rows(items, width):
    a = np.asarray(items, dtype=np.float64)
    return a.reshape(-1, width) if a.size else np.zeros((0, width))


def flatten_obstacles(cuboids=(), cylinders=()):
    """cuboids: iterable of (center[3], quaternion wxyz[4], dims[3]); cylinders: iterable of
    (center[3], quaternion wxyz[4], radius, height)  (the fields the reference reads off geometrout's Cuboid /
    Cylinder, load_test_dataset.py:105-116).  Returns (obstacle_config [no,10], cuboid_config [nb,10],
    cylinder_config [nc,9], num_cuboids, num_cylinders) exactly as TestDataset.fetch_data does (:188)."""
    cub = [np.concatenate([np.asarray(c, float).ravel(), np.asarray(q, float).ravel(), np.asarray(d, float).ravel()])
           for c, q, d in cuboids]
    cyl = [np.concatenate([np.asarray(c, float).ravel(), np.asarray(q, float).ravel(), [float(r)], [float(h)]])
           for c, q, r, h in cylinders]
    cub, cyl = _rows(cub, 10), _rows(cyl, 9)
    nb, nc = cub.shape[0], cyl.shape[0]
    if nb + nc == 0:
        raise ValueError("a scene needs at least one obstacle")
    # wxyz -> xyzw (np.roll(quats, -1, axis=1), :128 and :135)
    cuboid_config = np.concatenate([cub[:, 0:3], np.roll(cub[:, 3:7], -1, axis=1), cub[:, 7:10]], axis=1)
    cylinder_config = np.concatenate([cyl[:, 0:3], np.roll(cyl[:, 3:7], -1, axis=1), cyl[:, 7:8], cyl[:, 8:9]], axis=1)
    # cylinders enter the AABB guide as boxes with dims (r, r, h) -- radius, not diameter (:136-139)
    as_boxes = np.concatenate([cylinder_config[:, 0:7], cylinder_config[:, 7:8], cylinder_config[:, 7:8],
                               cylinder_config[:, 8:9]], axis=1)
    obstacle_config = np.concatenate([cuboid_config, as_boxes], axis=0)
    return obstacle_config, cuboid_config, cylinder_config, nb, nc


def sdf_primitives(cuboid_config, cylinder_config):
    """The same scene for the sphere / signed-distance family (edmp_b200.lib.SphereSDFGuide): boxes [nb,10] and
    cylinders [nc,9] = (xyz, quaternion xyzw, radius, height) pass through unchanged."""
    return (np.asarray(cuboid_config, dtype=np.float64).reshape(-1, 10),
            np.asarray(cylinder_config, dtype=np.float64).reshape(-1, 9))


def goal_volumes(guide, ik_goals):
    """[k,7] candidate goals -> float [k] total intersection volume of the goal configuration at t=0: one batched
    guide.cost call on [k,7,1] (infer_serial.py:119-121)."""
    import torch
    ik_goals = np.asarray(ik_goals, dtype=np.float64).reshape(-1, 7)
    k = ik_goals.shape[0]
    vol = guide.cost(torch.tensor(ik_goals[:, :, None], dtype=torch.float32), 0, batch_size=k)
    return vol.sum(dim=(1, 2)).cpu().numpy()


def select_goal(volumes, start, ik_goals, trust_region=GOAL_TRUST_REGION):
    """infer_serial.py:122-129: candidates with volume < min + trust_region, nearest to the start in joint space.
    Returns (goal [7], index into ik_goals)."""
    ik_goals = np.asarray(ik_goals, dtype=np.float64).reshape(-1, 7)
    volumes = np.asarray(volumes, dtype=np.float64).ravel()
    if volumes.shape[0] != ik_goals.shape[0] or ik_goals.shape[0] == 0:
        raise ValueError("need one volume per candidate goal and at least one candidate")
    keep = np.flatnonzero(volumes < volumes.min() + trust_region)
    dist = np.linalg.norm(ik_goals[keep] - np.asarray(start, dtype=np.float64)[None, :], axis=1)
    idx = int(keep[int(np.argmin(dist))])
    return ik_goals[idx], idx

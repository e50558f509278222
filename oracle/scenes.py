"""TEST INFRASTRUCTURE + bench inputs: synthetic planning problems.

MPiNets problem sets are a Google-Drive download (absent), so scenes are random boxes in the
Franka workspace in the reference's flattened format [no,10] = (xyz, quat xyzw, dims)
(datasets/load_test_dataset.py:76-189).  Pure numpy so both arms of the bench share it.
"""
import numpy as np

START = np.array([0.0, -0.5, 0.0, -2.0, 0.0, 1.6, 0.8])
GOAL = np.array([1.0, 0.3, -0.5, -1.5, 0.3, 2.0, 0.2])


def random_quats(n, rng):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def synthetic_scene(no=8, seed=1, rotated=True, cylinders=0):
    """[no,10] boxes: centres U([-0.2,-0.6,0],[0.8,0.6,0.8]) m, dims U(0.05,0.4) m.  The last
    ``cylinders`` entries are cylinders flattened the reference way, dims = (r, r, h)
    (load_test_dataset.py:136-139)."""
    rng = np.random.default_rng(seed)
    cfg = np.zeros((no, 10))
    cfg[:, 0:3] = rng.uniform([-0.2, -0.6, 0.0], [0.8, 0.6, 0.8], size=(no, 3))
    cfg[:, 3:7] = random_quats(no, rng) if rotated else np.array([0.0, 0.0, 0.0, 1.0])
    cfg[:, 7:10] = rng.uniform(0.05, 0.4, size=(no, 3))
    for i in range(no - cylinders, no):
        r = rng.uniform(0.03, 0.15)
        cfg[i, 7:10] = (r, r, rng.uniform(0.1, 0.5))
    return cfg


def tabletop_scene(seed=3, extra=3):
    """A mild, hand-placed scene around the START->GOAL sweep: a table slab plus a few boxes
    just outside the arm's interpolated sweep, so guidance acts on some waypoints only and the
    255-step chain stays well conditioned (used for the end-to-end <=1e-4 rad fixtures)."""
    rng = np.random.default_rng(seed)
    rows = [[0.30, 0.00, -0.30, 0, 0, 0, 1, 1.00, 1.20, 0.10],
            [0.80, -0.45, 0.30, 0, 0, 0, 1, 0.20, 0.25, 0.30],
            [0.15, -0.70, 0.50, 0, 0, 0, 1, 0.30, 0.10, 0.40],
            [-0.65, 0.30, 0.40, 0, 0, 0, 1, 0.15, 0.30, 0.30],
            [0.45, 0.95, 0.45, 0, 0, 0, 1, 0.35, 0.12, 0.50]]
    cfg = np.array(rows, dtype=np.float64)
    for i in range(1, min(1 + extra, len(rows))):           # yaw the side boxes a little
        ang = rng.uniform(-0.6, 0.6)
        cfg[i, 3:7] = (0.0, 0.0, np.sin(ang / 2), np.cos(ang / 2))
    return cfg


def gentle_x_T(rows, alpha_bar_T, seed=5, spread=0.05, start=START, goal=GOAL):
    """x_T = sqrt(alpha_bar_T) * (straight joint-space line + small noise): without a trained
    denoiser this keeps x_0 near the line instead of 3.6 x N(0,1) (far outside joint limits)."""
    rng = np.random.default_rng(seed)
    line = start[None, :, None] + (goal - start)[None, :, None] * np.linspace(0, 1, 50)[None, None, :]
    return np.sqrt(alpha_bar_T) * (line + spread * rng.normal(size=(rows, 7, 50)))

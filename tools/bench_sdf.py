"""Microbenchmark of the sphere / signed-distance guide kernels (BASELINE.json configs[4]): 50-waypoint Franka
trajectories against (a) box/cylinder primitives (cost + analytic gradient + clearance) and (b) a scene point cloud of
1k-64k points (nearest-point clearance).  Device time by CUDA events on the launching stream, after warm-up.
Reports ms, trajectories/s, GFLOP/s (cloud: 8 FLOP per (sphere, point) pair) and the
algorithmic HBM GB/s.   python tools/bench_sdf.py [out.txt]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import synthetic  # noqa: E402
from edmp_b200.lib import SphereSDFGuide  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    scene = synthetic.synthetic_scene(20, seed=1, rotated=True, cylinders=0)
    guide = SphereSDFGuide(scene, None, dev)
    print("# primitives: 20 boxes, 59 spheres x 50 waypoints per row; cost + gradient + clearance in one launch", file=out)
    for rows in (1024, 4096, 16384, 65536):
        q = torch.tensor(synthetic.gentle_x_T(rows, 1.0, seed=3, spread=0.2), dtype=torch.float32, device=dev)
        ms = timed(lambda: guide.evaluate(q))
        pairs = rows * 50 * 59 * 20
        gb = rows * 50 * 7 * 4 * 2 + rows * 50 * 4 + rows * 4          # q in, grad out, clearance, cost
        print("primitives rows %6d: %8.3f ms  %10.0f traj/s  %7.1f G (sphere,primitive) pairs/s  %6.1f GB/s algorithmic"
              % (rows, ms, rows / ms * 1e3, pairs / ms / 1e6, gb / ms / 1e6), file=out)
    print("# point cloud: clearance = min over (sphere, point) of |c - p| - r; 8 FLOP per pair", file=out)
    for rows, npts in ((1024, 1024), (1024, 16384), (1024, 65536), (8192, 4096), (65536, 1024)):
        q = torch.tensor(synthetic.gentle_x_T(rows, 1.0, seed=3, spread=0.2), dtype=torch.float32, device=dev)
        pts = torch.tensor(rng.uniform([-0.3, -0.7, 0.0], [0.9, 0.7, 0.9], size=(npts, 3)), dtype=torch.float32, device=dev)
        ms = timed(lambda: guide.cloud_clearance(pts, q), iters=3)
        pairs = rows * 50 * 59 * npts
        gb = rows * 50 * 7 * 4 + rows * npts * 16 + rows * 50 * 4       # every row streams the cloud through L2
        print("cloud rows %6d points %6d: %9.3f ms  %10.0f traj/s  %8.1f GFLOP/s  %7.1f GB/s (cloud re-read per row)"
              % (rows, npts, ms, rows / ms * 1e3, 8.0 * pairs / ms / 1e6, gb / ms / 1e6), file=out)


if __name__ == "__main__":
    main()

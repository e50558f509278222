"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/sdf.npz: signed distances of random points to random cuboids and
cylinders computed by the UNMODIFIED reference (mpinets/geometry.py TorchCuboids / TorchCylinders, imported from
/root/reference under a geometrout stub), for pinning oracle/sdf_oracle.py and the CUDA kernel.  Run in the build
container:  python -m oracle.make_golden_sdf"""
import os

import numpy as np
import torch

from . import sdf_oracle as so


def random_scene(rng, nb, nc, yaw_only=False):
    def quats(n):
        if yaw_only:
            a = rng.uniform(-np.pi, np.pi, n)
            return np.stack([np.zeros(n), np.zeros(n), np.sin(a / 2), np.cos(a / 2)], axis=1)
        q = rng.normal(size=(n, 4))
        return q / np.linalg.norm(q, axis=1, keepdims=True)
    boxes = np.zeros((nb, 10))
    boxes[:, 0:3] = rng.uniform([-0.3, -0.7, 0.0], [0.9, 0.7, 0.9], size=(nb, 3))
    boxes[:, 3:7] = quats(nb)
    boxes[:, 7:10] = rng.uniform(0.05, 0.4, size=(nb, 3))
    cyls = np.zeros((nc, 9))
    cyls[:, 0:3] = rng.uniform([-0.3, -0.7, 0.0], [0.9, 0.7, 0.9], size=(nc, 3))
    cyls[:, 3:7] = quats(nc)
    cyls[:, 7] = rng.uniform(0.03, 0.15, nc)
    cyls[:, 8] = rng.uniform(0.1, 0.5, nc)
    return boxes, cyls


def main():
    geo = so.load_reference_geometry()
    out = {}
    for name, yaw in (("general", False), ("yaw", True)):
        rng = np.random.default_rng(11 if yaw else 7)
        boxes, cyls = random_scene(rng, 9, 4, yaw_only=yaw)
        pts = rng.uniform([-0.6, -0.9, -0.2], [1.1, 0.9, 1.2], size=(1, 800, 3))
        wxyz = lambda a: a[:, [3, 0, 1, 2]]     # obstacle_config rows are xyzw, mpinets wants wxyz (geometry.py:60-63)
        C = geo.TorchCuboids(torch.tensor(boxes[None, :, 0:3]), torch.tensor(boxes[None, :, 7:10]),
                             torch.tensor(wxyz(boxes[:, 3:7])[None]))
        Y = geo.TorchCylinders(torch.tensor(cyls[None, :, 0:3]), torch.tensor(cyls[None, :, 7:8]),
                               torch.tensor(cyls[None, :, 8:9]), torch.tensor(wxyz(cyls[:, 3:7])[None]))
        out[name + "/boxes"], out[name + "/cylinders"], out[name + "/points"] = boxes, cyls, pts[0]
        out[name + "/cuboid_sdf"] = C.sdf(torch.tensor(pts))[0].numpy()
        out[name + "/cylinder_sdf"] = Y.sdf(torch.tensor(pts))[0].numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sdf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

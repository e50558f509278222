"""Synthetic planning problems and seeded weights (host side, numpy/torch-CPU only).

The reference's inputs are downloads that are not part of either repository: the MpiNets problem sets
(datasets/*_solvable_problems.pkl, reference README.md:53-58, .gitignore:1-4) and the trained
TemporalUNetModel255_N50 weights.  `infer_serial.py` (dataset_type: 'synthetic') and `bench.py` therefore run on
problems generated here: obstacles in the reference's flattened format [no,10] = (xyz, quaternion xyzw, dims)
(datasets/load_test_dataset.py:76-189, cylinders flattened to (r, r, h) boxes :136-139), a start configuration,
candidate goal configurations (the reference gets them from ikfast, load_test_dataset.py:176-186), and a seeded
state_dict with the reference's keys and shapes (diffusion/models/temporalunet.py:11-36).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

START = np.array([0.0, -0.5, 0.0, -2.0, 0.0, 1.6, 0.8])
GOAL = np.array([1.0, 0.3, -0.5, -1.5, 0.3, 2.0, 0.2])
DIMS = (32, 64, 128, 256, 512, 512)
TIME_DIM = 32


def synthetic_scene(no=8, seed=1, rotated=True, cylinders=0):
    """[no,10] boxes: centres U([-0.2,-0.6,0],[0.8,0.6,0.8]) m, dims U(0.05,0.4) m; the last ``cylinders`` entries are
    cylinders flattened the reference way, dims = (r, r, h)."""
    rng = np.random.default_rng(seed)
    cfg = np.zeros((no, 10))
    cfg[:, 0:3] = rng.uniform([-0.2, -0.6, 0.0], [0.8, 0.6, 0.8], size=(no, 3))
    if rotated:
        q = rng.normal(size=(no, 4))
        cfg[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    else:
        cfg[:, 3:7] = np.array([0.0, 0.0, 0.0, 1.0])
    cfg[:, 7:10] = rng.uniform(0.05, 0.4, size=(no, 3))
    for i in range(no - cylinders, no):
        r = rng.uniform(0.03, 0.15)
        cfg[i, 7:10] = (r, r, rng.uniform(0.1, 0.5))
    return cfg


def tabletop_scene(seed=3, extra=3):
    """A table slab plus a few boxes just outside the arm's START -> GOAL sweep."""
    rng = np.random.default_rng(seed)
    rows = [[0.30, 0.00, -0.30, 0, 0, 0, 1, 1.00, 1.20, 0.10],
            [0.80, -0.45, 0.30, 0, 0, 0, 1, 0.20, 0.25, 0.30],
            [0.15, -0.70, 0.50, 0, 0, 0, 1, 0.30, 0.10, 0.40],
            [-0.65, 0.30, 0.40, 0, 0, 0, 1, 0.15, 0.30, 0.30],
            [0.45, 0.95, 0.45, 0, 0, 0, 1, 0.35, 0.12, 0.50]]
    cfg = np.array(rows, dtype=np.float64)
    for i in range(1, min(1 + extra, len(rows))):
        ang = rng.uniform(-0.6, 0.6)
        cfg[i, 3:7] = (0.0, 0.0, np.sin(ang / 2), np.cos(ang / 2))
    return cfg


def goal_candidates(k=8, seed=7, goal=GOAL, spread=0.15):
    """[k,7] candidate goal configurations around ``goal`` (stand-in for the ikfast solutions the reference samples
    for the target pose); the entry point filters them with the guide's t=0 cost (infer_serial.py:119-129)."""
    rng = np.random.default_rng(seed)
    out = goal[None, :] + spread * rng.normal(size=(k, 7))
    out[0] = goal
    return out


def gentle_x_T(rows, alpha_bar_T, seed=5, spread=0.05, start=START, goal=GOAL):
    """x_T = sqrt(alpha_bar_T) * (straight joint-space line + small noise): with untrained (seeded) weights this keeps
    x_0 near the line instead of 3.6 x N(0,1), far outside the joint limits."""
    rng = np.random.default_rng(seed)
    line = start[None, :, None] + (goal - start)[None, :, None] * np.linspace(0, 1, 50)[None, None, :]
    return np.sqrt(alpha_bar_T) * (line + spread * rng.normal(size=(rows, 7, 50)))


def alpha_bar_T(T=255, thresh=0.02):
    """alpha_bar at the last step of the linear-beta schedule (diffusion/diffusion.py:13-16,:47)."""
    beta = np.linspace(0.0, thresh, T + 1)[1:]
    return float(np.prod(1.0 - beta))


def _res_block(prefix, cin, cout, out):
    for b, c in ((0, cin), (1, cout)):
        out.append((prefix + ".blocks.%d.block.0.weight" % b, (cout, c, 5), "w"))
        out.append((prefix + ".blocks.%d.block.0.bias" % b, (cout,), "b%d" % (c * 5)))
        out.append((prefix + ".blocks.%d.block.2.weight" % b, (cout,), "gamma"))
        out.append((prefix + ".blocks.%d.block.2.bias" % b, (cout,), "beta"))
    out.append((prefix + ".time_mlp.time_mlp.1.weight", (cout, TIME_DIM), "w"))
    out.append((prefix + ".time_mlp.time_mlp.1.bias", (cout,), "b%d" % TIME_DIM))
    if cin != cout:
        out.append((prefix + ".residual_conv.weight", (cout, cin, 1), "w"))
        out.append((prefix + ".residual_conv.bias", (cout,), "b%d" % cin))


def key_table(input_dim=7, dims=DIMS):
    """[(key, shape, kind)] in the reference's state_dict order."""
    d = [input_dim, *dims]
    out = [("time_embedding.time_mlp.1.weight", (4 * TIME_DIM, TIME_DIM), "w"),
           ("time_embedding.time_mlp.1.bias", (4 * TIME_DIM,), "b%d" % TIME_DIM),
           ("time_embedding.time_mlp.3.weight", (TIME_DIM, 4 * TIME_DIM), "w"),
           ("time_embedding.time_mlp.3.bias", (TIME_DIM,), "b%d" % (4 * TIME_DIM))]
    n_down = len(d) - 1
    for i in range(n_down):
        _res_block("down_samplers.%d.down.0" % i, d[i], d[i + 1], out)
        _res_block("down_samplers.%d.down.1" % i, d[i + 1], d[i + 1], out)
        if i != n_down - 1:
            out.append(("down_samplers.%d.down.3.weight" % i, (d[i + 1], d[i + 1], 3), "w"))
            out.append(("down_samplers.%d.down.3.bias" % i, (d[i + 1],), "b%d" % (d[i + 1] * 3)))
    _res_block("middle_block.middle.0", d[-1], d[-1], out)
    _res_block("middle_block.middle.2", d[-1], d[-1], out)
    for n, i in enumerate(range(len(d) - 1, 1, -1)):
        _res_block("up_samplers.%d.up.0" % n, 2 * d[i], d[i - 1], out)
        _res_block("up_samplers.%d.up.1" % n, d[i - 1], d[i - 1], out)
        out.append(("up_samplers.%d.up.3.weight" % n, (d[i - 1], d[i - 1], 4), "w"))
        out.append(("up_samplers.%d.up.3.bias" % n, (d[i - 1],), "b%d" % (d[i - 1] * 4)))
    out.append(("final_conv.0.block.0.weight", (d[1], d[1], 5), "w"))
    out.append(("final_conv.0.block.0.bias", (d[1],), "b%d" % (d[1] * 5)))
    out.append(("final_conv.0.block.2.weight", (d[1],), "gamma"))
    out.append(("final_conv.0.block.2.bias", (d[1],), "beta"))
    out.append(("final_conv.1.weight", (input_dim, d[1], 1), "w"))
    out.append(("final_conv.1.bias", (input_dim,), "b%d" % d[1]))
    return out


def seeded_state_dict(seed=0, gain=1.0, input_dim=7, dims=DIMS, final_gain=1.0):
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) * gain for conv/linear weights and biases (torch's default-init bound),
    GroupNorm gamma = 1 + 0.2 U(-1,1), beta = 0.2 U(-1,1).  ``final_gain`` scales the last 1x1 conv: an untrained
    net's eps is an arbitrary drift, and 0.2 keeps the 255-step chain in the joint limits' interior."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in key_table(input_dim, dims):
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[key] = u * (gain / math.sqrt(fan_in))
        elif kind[0] == "b" and kind[1:].isdigit():
            sd[key] = u * (1.0 / math.sqrt(int(kind[1:])))
        elif kind == "gamma":
            sd[key] = 1.0 + 0.2 * u
        else:
            sd[key] = 0.2 * u
    if final_gain != 1.0:
        sd["final_conv.1.weight"] = sd["final_conv.1.weight"] * final_gain
        sd["final_conv.1.bias"] = sd["final_conv.1.bias"] * final_gain
    return sd

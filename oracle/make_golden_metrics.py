"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/metrics.npz: path lengths, SPARC smoothness and end-effector
transforms computed by the UNMODIFIED reference (lib/metrics.py MetricsCalculator on lib/guide.py
IntersectionVolumeGuide, imported from /root/reference through oracle/ref_shim.py), for pinning
oracle/metrics_oracle.py and the CUDA kernels.  Run in the build container:  python -m oracle.make_golden_metrics"""
import importlib.util
import os

import numpy as np
import torch

from . import ref_shim


def trajectories():
    """[R, 7, 50]: sampled trajectories of the sampler fixture, smooth interpolations with sinusoidal detours of
    growing frequency, a noisy one, and a constant one (all-zero movement, lib/metrics.py:86-88)."""
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "..", "tests", "golden", "sampler_iv.npz"))
    rows = [gold["final"][i] for i in range(4)]
    start = np.array([0, -0.5, 0, -2.0, 0, 1.6, 0.8])
    goal = np.array([1.0, 0.3, -0.5, -1.5, 0.3, 2.0, 0.2])
    s = np.linspace(0.0, 1.0, 50)
    rng = np.random.default_rng(5)
    for k in range(4):
        smooth = 3 * s ** 2 - 2 * s ** 3
        x = start[:, None] + (goal - start)[:, None] * smooth[None, :]
        x = x + 0.15 * np.sin(np.pi * (k + 1) * s)[None, :] * rng.normal(size=(7, 1))
        rows.append(x)
    rows.append(rows[4] + 0.01 * rng.normal(size=(7, 50)))
    rows.append(np.repeat(start[:, None], 50, axis=1))
    return np.stack(rows)


def main():
    ns = ref_shim.load_reference()
    spec = importlib.util.spec_from_file_location("_edmp_ref_metrics", os.path.join(ref_shim.REF_ROOT, "lib", "metrics.py"))
    ref_metrics = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_metrics)
    guide = ns.IntersectionVolumeGuide(obstacle_config=np.zeros((1, 10)), device="cpu", guide_cfgs={}, batch_size=1)
    calc = ref_metrics.MetricsCalculator(guide)
    traj = trajectories()
    dts = np.array([0.02, 0.05, 0.1])      # fs = 50, 20, 10 Hz: the last one selects mirrored bins too (f <= fc everywhere)
    out = {"traj": traj, "dts": dts}
    R = traj.shape[0]
    paths = np.zeros((R, 2))
    sal = np.zeros((len(dts), R, 2))
    for r in range(R):
        paths[r] = calc.path_length_metric(traj[r])
        for i, dt in enumerate(dts):
            js, es = calc.smoothness_metric(traj[r], float(dt))
            sal[i, r, 0] = js[0]
            sal[i, r, 1] = es[0]
            if r == 0 and i == 0:
                out["f"] = js[1][0]
                out["Mf_joint"] = js[1][1]
                out["Mf_ee"] = np.asarray(es[1][1], dtype=np.float64)
                out["fsel_joint"] = js[2][0]
    out["path_lengths"] = paths
    out["sparc"] = sal
    q = torch.tensor(traj, dtype=torch.float32).permute(0, 2, 1)
    out["ee_transforms"] = guide.get_end_effector_transform(q).numpy()
    # the docstring example of the reference's sparc (lib/metrics.py:79-84): -1.41403
    t = np.arange(-1, 1, 0.01)
    move = np.exp(-5 * t ** 2)
    out["example_move"] = move
    out["example_sal"] = np.array(calc.sparc(move, fs=100.)[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), {k: np.asarray(v).shape for k, v in out.items()})
    print("sparc", sal[0], "example", out["example_sal"])


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""infer_serial.py -c <benchmark cfg>  --  the reference's entry point, kept as the way to run the sampler.

Same command line (`-c/--cfg_path`, reference infer_serial.py:19), same benchmark YAML (`guide`, `dataset`, `model`,
`general` sections, benchmark/cfgs/cfg1.yaml:1-24) and per-guide YAML plugins (guides/cfgs/guide<N>.yaml), same
sequence per planning problem (reference infer_serial.py:95-167):

    guide tables -> IntersectionVolumeGuide(scene) -> goal filter on the guide's t=0 cost -> denoise_guided
    -> best-of-ensemble -> success check

Everything between those calls runs in libedmp_b200.so on the GPU (edmp_b200.Diffusion / TemporalUNet /
IntersectionVolumeGuide are thin ctypes shims); this file is host glue only.  What is NOT rebuilt (SURVEY.md
section 8f): the MpiNets problem-set loader with its ikfast goal sampling and the PyBullet rollout.  With
`dataset_type: 'synthetic'` the problems come from edmp_b200.synthetic and success is the guide's own swept-volume
test of the chosen trajectory; any other dataset_type asks for the reference's `datasets.load_test_dataset`
(put the reference checkout and its downloads on PYTHONPATH / in ./datasets).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from edmp_b200 import (Diffusion, IntersectionVolumeGuide, TemporalUNet, YamlConfig, build_guide_cfgs,  # noqa: E402
                       load_guide_hparams, synthetic)
from edmp_b200 import scene as front_end  # noqa: E402
from edmp_b200.lib import MetricsCalculator, RobotEnvironment  # noqa: E402

DIMS = (32, 64, 128, 256, 512, 512)        # reference infer_serial.py:50
GOAL_TRUST_REGION = front_end.GOAL_TRUST_REGION   # reference infer_serial.py:125 (hard-coded there too)


SYNTHETIC_SCENES_PER_TYPE = 2   # size of the synthetic problem set per scene type (what "all scenes" means for it)


class _EveryType(dict):
    def __missing__(self, key):
        return SYNTHETIC_SCENES_PER_TYPE


class SyntheticProblems:
    """Stand-in for datasets.load_test_dataset.TestDataset: fetch_data(scene_num, scene_type) returns the fields the
    entry point uses (obstacle_config [no,10], start [7], all_ik_goals [k,7]); data_nums[scene_type] = problems per type
    like the reference loader's attribute (infer_serial.py:98)."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self.data_nums = _EveryType()

    def fetch_data(self, scene_num, scene_type="tabletop"):
        if scene_type == "tabletop":
            scene = synthetic.tabletop_scene(seed=3 + scene_num)
        else:
            scene = synthetic.synthetic_scene(no=20, seed=1 + self.seed + scene_num, rotated=True, cylinders=4)
        return scene, synthetic.START.copy(), synthetic.goal_candidates(8, seed=7 + scene_num)


def load_problems(cfg):
    ds = cfg["dataset"]
    if ds["dataset_type"] == "synthetic":
        return SyntheticProblems(ds.get("seed", 0)), ds.get("scene_types", ["tabletop"])
    try:
        from datasets.load_test_dataset import TestDataset   # the reference's loader (needs its downloads)
    except Exception as e:   # noqa: BLE001
        raise SystemExit("infer_serial.py: dataset_type %r needs the reference's datasets.load_test_dataset and the "
                         "MpiNets problem sets (not part of this build, SURVEY.md section 8f-1): %s" %
                         (ds["dataset_type"], e))
    real = TestDataset(ds["dataset_type"], d_path=ds["path"])

    class _Adapter:
        data_nums = real.data_nums

        def fetch_data(self, scene_num, scene_type):
            obstacle_config, _, _, _, _, start, goals = real.fetch_data(scene_num=scene_num, scene_type=scene_type)
            return obstacle_config, start, goals

    return _Adapter(), ds["scene_types"]


def scenes_per_type(cfg, problems, scene_type):
    """How many problems of `scene_type` to run.  The reference always walks the whole set (infer_serial.py:98 loops
    over dataset.data_nums[scene_type]; its cfg1.yaml says `num_scenes_per_type: -1  # -1 implies all scenes`), so a
    missing key or a value <= 0 means ALL; a positive value caps the walk."""
    total = int(problems.data_nums[scene_type])
    want = cfg["dataset"].get("num_scenes_per_type", -1)
    if want is None or int(want) <= 0:
        return total
    return min(int(want), total)


def load_model(cfg, device, precision):
    m = cfg["model"]
    name = os.path.join(m["model_dir"], "TemporalUNetModel255_N50")
    seed = m.get("synthetic_weights_seed")
    have = os.path.exists(os.path.join(name, "weights_latest.pt")) and os.path.exists(os.path.join(name, "losses.npy"))
    if not have and seed is None:
        # the reference asks on stdin and exits (infer_serial.py:46-49); a batch entry point just fails
        raise SystemExit("infer_serial.py: no checkpoint under %s (weights_latest.pt + losses.npy) and no "
                         "model.synthetic_weights_seed in the cfg" % name)
    if not have:
        import tempfile
        name = os.path.join(tempfile.mkdtemp(prefix="edmp_model_"), "TemporalUNetModel255_N50")
    model = TemporalUNet(model_name=name, input_dim=m["num_channels"], time_dim=32, dims=DIMS, device=device,
                         precision=precision)
    if not have:
        model.load_state_dict(synthetic.seeded_state_dict(int(seed), final_gain=0.2))
        print("No checkpoint found: seeded synthetic weights (seed %d)" % int(seed))
    return model


def pick_goal(guide, start, ik_goals):
    """Goal filter of the reference (infer_serial.py:119-129), see edmp_b200.scene."""
    vol = front_end.goal_volumes(guide, ik_goals)
    goal, _ = front_end.select_goal(vol, start, ik_goals, GOAL_TRUST_REGION)
    return goal, vol


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--cfg_path", default="./benchmark/cfgs/cfg1.yaml")
    ap.add_argument("--precision", default=os.environ.get("EDMP_PRECISION", "f16x3"),
                    help="arithmetic mode of the UNet contractions (f16x3, tf32x3, fp32 are parity grade)")
    ap.add_argument("--noise", default="numpy", choices=["numpy", "philox"],
                    help="numpy: the reference's np.random draws in the reference's order; philox: on the device")
    args = ap.parse_args(argv)

    cfg = YamlConfig(args.cfg_path)
    device = cfg["model"]["device"]
    guides = cfg["guide"]["guides"]
    bpg = cfg["guide"]["batch_size_per_guide"]
    hp = load_guide_hparams(guides, cfg["guide"]["guide_path"])
    guide_cfgs = build_guide_cfgs(hp, bpg, T=cfg["model"]["T"])                  # infer_serial.py:59-91
    total = guide_cfgs["total_batch_size"]

    problems, scene_types = load_problems(cfg)
    env = RobotEnvironment(gui=cfg["general"]["gui"])
    diffuser = Diffusion(T=cfg["model"]["T"], device=device)
    diffuser.noise_mode = args.noise
    model = load_model(cfg, device, args.precision)
    save_dir = cfg["general"].get("save_dir")

    results, n_success = [], 0
    for scene_type in scene_types:
        for scene_num in range(scenes_per_type(cfg, problems, scene_type)):
            t_plan = time.time()
            obstacle_config, start, ik_goals = problems.fetch_data(scene_num, scene_type)
            guide = IntersectionVolumeGuide(obstacle_config=obstacle_config, device=device, guide_cfgs=guide_cfgs,
                                            batch_size=total)
            _ = MetricsCalculator(guide)
            t0 = time.time()
            goal, goal_vol = pick_goal(guide, start, np.asarray(ik_goals, dtype=np.float64))
            t_ik = time.time() - t0
            t0 = time.time()
            trajectories = diffuser.denoise_guided(model=model, guide=guide, batch_size=total,
                                                   traj_len=cfg["model"]["traj_len"],
                                                   num_channels=cfg["model"]["num_channels"], condition=True,
                                                   benchmarking=True, start=start, goal=goal,
                                                   guidance_schedule=guide_cfgs["guidance_schedule"])
            t_denoise = time.time() - t0
            best = guide.choose_best_trajectory(start, goal, trajectories)        # lib/guide.py:637-653
            env.clear_obstacles()
            env.attach_guide(guide, start, goal)
            success = env.benchmark_trajectory(best)
            n_success += success
            print("%s scene %d: %d rows (%d guides x %d), goal filter %.3f s, denoiser %.3f s, planning %.3f s, "
                  "success %d" % (scene_type, scene_num, total, len(guides), bpg, t_ik, t_denoise,
                                  time.time() - t_plan, success))
            results.append({"scene_type": scene_type, "scene_num": scene_num, "trajectory": best,
                            "trajectories": trajectories, "goal": goal, "start": start, "success": success,
                            "goal_volumes": goal_vol, "denoise_s": t_denoise})
            if save_dir:
                os.makedirs(save_dir, exist_ok=True)
                np.save(os.path.join(save_dir, "%s_%d_best.npy" % (scene_type, scene_num)), best)
    print("Success: %d / %d" % (n_success, len(results)))
    return results


if __name__ == "__main__":
    main()

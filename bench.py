#!/usr/bin/env python
"""bench.py -- guided-diffusion trajectory sampling throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is ONE pass of the hot path over one batch: the full 255-step guided reverse diffusion
(TemporalUNet + posterior + guide gradient/update) of every trajectory row of the batch, followed
by the per-row best-of-ensemble cost.  Workload (BASELINE.json configs[1]): the
"batch=8192 trajectories, guides [1..10]" ensemble fits one GPU, so at N = 1 it runs whole: the first
ten shipped guides [1,2,3,4,5,9,10,11,12,13] x 819 rows = 8190 trajectories per GPU, 20 synthetic
obstacles, seeded random weights.  Weak scaling: every rank runs its own ensemble of that size
(ensembles are rank-local), one NCCL all-gather of the per-row final costs.  `--rows-per-guide 102`
gives the 1/8 shard (1020 rows/GPU) of the strong-scaling reading of the same config.

Prints ONE JSON line (rank 0).  `value` = trajectories/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host x_T in, trajectories + costs out).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GUIDES = [1, 2, 3, 4, 5, 9, 10, 11, 12, 13]
ROWS_PER_GUIDE = 819
N_OBSTACLES = 20
USEFUL_GFLOP_PER_ROW_STEP = 0.1222     # 61,096,192 non-padding MACs
METRIC = "trajectories/sec (255-step, 50x7-DoF, guided ensemble)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="edmp_clocks_", suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                out["sm_max_mhz"] = float(f[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def build_workload(seed_offset=0, rows_per_guide=ROWS_PER_GUIDE, guides=GUIDES):
    from edmp_b200 import build_guide_cfgs, load_guide_hparams
    from oracle import sampler_oracle as so, scenes          # input generators only (shared by both arms)
    hp = load_guide_hparams(guides, os.path.join(ROOT, "guides") + "/")
    cfgs = build_guide_cfgs(hp, rows_per_guide)
    scene = scenes.synthetic_scene(N_OBSTACLES, seed=1, rotated=True, cylinders=4)
    _, _, abar = so.schedule()
    rows = cfgs["total_batch_size"]
    # x_T near the straight joint-space line keeps the untrained net's chain finite
    x_T = scenes.gentle_x_T(rows, abar[-1], seed=100 + seed_offset)
    return cfgs, scene, x_T, scenes.START.copy(), scenes.GOAL.copy()


def synthetic_state_dict():
    from oracle import weights                                # seeded random weights (real ones are a download)
    return weights.seeded_state_dict(0, final_gain=0.2)


# -----------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port; /root/reference does not exist on the GPU box)
# -----------------------------------------------------------------------------------------------------
def cpu_sample(rows_per_guide, sd, guides=GUIDES, seed=0):
    import torch
    from oracle import sampler_oracle as so, guide_oracle as go
    cfgs, scene, x_T, start, goal = build_workload(seed, rows_per_guide, guides)
    rows = cfgs["total_batch_size"]
    rng = np.random.default_rng(7 + seed)
    noise = [rng.normal(size=(rows, 7, 50)) for _ in range(255)]
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        out = so.denoise_guided(sd, scene, cfgs, start, goal, x_T, noise, gradient="autograd")
        go.final_sv_costs(np.nan_to_num(out), start, goal, scene)
    return rows, time.perf_counter() - t0


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = synthetic_state_dict()
    cores = torch.get_num_threads()
    for _ in range(args.warmup):
        cpu_sample(1, sd)
    times = []
    rows = 0
    for k in range(args.steps):
        rows, dt = cpu_sample(1, sd, seed=k)
        times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = rows / (ms / 1000.0)
    sample = "%d rows (10 guides x 1) x 255 steps, %d obstacles per step; oracle port (torch CPU fp32 UNet + " \
             "autograd guide), %d threads" % (rows, N_OBSTACLES, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, note="CPU arm runs a bounded sample of the same workload",
                                      rows_per_guide=args.rows_per_guide),
            "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, note=None, precision="fp32", rows_per_guide=None):
    rpg = rows_per_guide or ROWS_PER_GUIDE
    rows = len(GUIDES) * rpg
    cfg = {"workload": "configs[1]: guided ensemble, guides %s x %d rows = %d rows/GPU (%d total), %d obstacles, "
                       "T=255, horizon 50, 7 DoF" % (GUIDES, rpg, rows, rows * n_gpus, N_OBSTACLES),
           "rows_per_gpu": rows, "n_guides": len(GUIDES), "rows_per_guide": rpg,
           "obstacles": N_OBSTACLES, "precision_mode": precision, "parallelism": "dp%d (ensembles rank-local)" % n_gpus,
           "l2": "inputs larger than L2: one UNet forward streams ~57 MB of weight tiles and ~0.25 MB of "
                 "activations per row (%.1f GB at this batch) through the 126 MB L2; no explicit flush" % (rows * 0.25e-3)}
    if note:
        cfg["note"] = note
    return cfg


# -----------------------------------------------------------------------------------------------------
# GPU arm
# -----------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from edmp_b200 import Diffusion, IntersectionVolumeGuide, TemporalUNet, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = _lib.load()

    sd = synthetic_state_dict()
    model = TemporalUNet(os.path.join(tempfile.mkdtemp(), "TemporalUNetModel255_N50"), 7, 32, dev,
                         dims=(32, 64, 128, 256, 512, 512), precision=args.precision)
    model.load_state_dict(sd)
    cfgs, scene, x_T, start, goal = build_workload(seed_offset=rank, rows_per_guide=args.rows_per_guide)
    rows = cfgs["total_batch_size"]
    guide = IntersectionVolumeGuide(scene, dev, cfgs, rows)
    diff = Diffusion(255, dev)

    x0 = torch.tensor(x_T, dtype=torch.float64, device=dev)
    x = torch.empty_like(x0)
    gathered = [torch.empty(rows, device=dev, dtype=torch.float32) for _ in range(world)] if world > 1 else None

    def one_pass(seed):
        x.copy_(x0)
        cost = diff.run_steps(model, guide, x, start, goal, 255, 0, noise=None, seed=seed,
                              guidance_schedule=cfgs["guidance_schedule"], want_cost=True)
        if world > 1:
            dist.all_gather(gathered, cost)        # the single collective: per-row final costs
            allc = torch.stack(gathered)
        else:
            allc = cost[None]
        return torch.argmin(torch.nan_to_num(allc, nan=float("inf")), dim=1)   # best row per ensemble

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        one_pass(1000 + w)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        one_pass(k)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = diff.last_launches * args.steps
    tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    value = rows * world / (ms_per_step / 1000.0)

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------------
    xh = torch.tensor(x_T, dtype=torch.float64).pin_memory()
    xh_work = torch.empty_like(xh).pin_memory()
    ch = torch.empty(rows, dtype=torch.float32).pin_memory()
    s_arr, s_ptr = _lib.host_f64(start)
    g_arr, g_ptr = _lib.host_f64(goal)

    def e2e_pass(seed):
        xh_work.copy_(xh)
        with torch.cuda.device(dev):
            _lib.check(lib.edmp_sample_guided_host(
                diff._sampler(rows), model.engine(rows),
                guide.scene_handle(rows=rows, guidance_schedule=cfgs["guidance_schedule"]),
                ctypes.c_void_p(xh_work.data_ptr()), s_ptr, g_ptr, ctypes.c_uint64(seed), rows,
                ctypes.c_void_p(ch.data_ptr()), _lib.stream_ptr()), "edmp_sample_guided_host")
        return int(np.argmin(np.nan_to_num(ch.numpy(), nan=np.inf)))

    e2e_pass(5)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_pass(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rows * world * args.steps / float(te.item())
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: per-op CUDA-event times of UNet forwards -----------------
    roofline = None
    unet_summary = None
    if rank == 0:
        n_ops = lib.edmp_unet_launches_per_forward(model.engine(rows))
        ms = np.zeros(n_ops, dtype=np.float32)
        macs = np.zeros(n_ops, dtype=np.float64)
        xf = torch.randn(rows, 7, 50, device=dev)
        eps = torch.empty_like(xf)
        _lib.check(lib.edmp_unet_profile(model.engine(rows), ctypes.c_void_p(xf.data_ptr()), 128, rows, 10,
                                         ms.ctypes.data_as(ctypes.c_void_p), macs.ctypes.data_as(ctypes.c_void_p),
                                         ctypes.c_void_p(eps.data_ptr()), _lib.stream_ptr()), "edmp_unet_profile")
        names = [lib.edmp_unet_op_name(model.engine(rows), i).decode() for i in range(n_ops)]
        kernels = [lib.edmp_unet_op_kernel(model.engine(rows), i).decode() for i in range(n_ops)]
        # dominant kernel = the kernel (all its launches of one forward together) with the largest time share
        share = {}
        for k, m in zip(kernels, ms):
            share[k] = share.get(k, 0.0) + float(m)
        dom = max(share, key=share.get)
        sel = np.array([k == dom for k in kernels])
        top = int(np.argmax(np.where(sel, ms, 0.0)))
        peaks = load_peaks()
        prec = args.precision
        peak = peaks["bf16_tflops_sustained"] * (0.5 if prec in ("fp32", "tf32", "tf32x3") else 1.0)
        n_l = int(sel.sum())
        achieved = 2.0 * macs[sel].sum() / (ms[sel].sum() * 1e-3) / 1e12     # useful FLOPs of its launches / their time
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(prec, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16 sustained%s (%s)" %
                                   (" x 0.5 (tf32-class operands)" if peak != peaks["bf16_tflops_sustained"] else "",
                                    peaks["source"]),
                    "launches_per_forward": n_l, "kernel_ms": float(ms[sel].sum() / n_l),
                    "kernel_share_of_unet": float(ms[sel].sum() / ms.sum()),
                    "useful_flops_per_launch": float(2.0 * macs[sel].sum() / n_l),
                    "note": "useful (non-padding, un-split) FLOPs; the hi/lo operand split issues 3x as many MMA FLOPs",
                    "slowest_launch": {"op": names[top], "ms": float(ms[top]),
                                       "achieved": float(2.0 * macs[top] / (ms[top] * 1e-3) / 1e12)},
                    "by_kernel_ms": {k: round(v, 4) for k, v in share.items()}}
        if args.ops_out:
            with open(args.ops_out, "w") as f:
                for nm, kn, m, mc in zip(names, kernels, ms, macs):
                    f.write("%-44s %-14s %9.1f us %8.2f useful TFLOP/s\n" % (nm, kn, m * 1e3, 2 * mc / (m * 1e-3) / 1e12 if m > 0 else 0))
        unet_summary = {"ms_per_forward": float(ms.sum()), "useful_tflops": float(2.0 * macs.sum() / (ms.sum() * 1e-3) / 1e12),
                        "launches": int(n_ops)}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = torch.get_num_threads()
        crow, cdt = cpu_sample(2, sd)
        cpu_baseline = {"value": crow / cdt, "unit": "trajectories/s", "cores": cores, "kind": "port",
                        "sample": "%d rows (10 guides x 2) x 255 steps, %d obstacles, %.1f s; oracle port "
                                  "(torch CPU fp32 UNet + autograd guide)" % (crow, N_OBSTACLES, cdt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else args.precision, "data": "synthetic",
                "config": workload_config(world, precision=args.precision, rows_per_guide=args.rows_per_guide),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": rows * 350 * 8,
                        "d2h_bytes_per_step": rows * 350 * 8 + rows * 4},
                "gpu_launches": int(launches),
                "roofline": roofline, "unet": unet_summary, "cpu_baseline": cpu_baseline,
                "useful_tflops_whole_job": value * 255 * USEFUL_GFLOP_PER_ROW_STEP / 1e3}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("EDMP_PRECISION", "f16x3"),
                    help="f16x3 (tcgen05, IEEE-half hi/lo split, parity grade, default) | tf32x3 | fp32 (CUDA cores) | "
                         "bf16x3 / f16 / bf16 / tf32 (not parity grade)")
    ap.add_argument("--rows-per-guide", type=int, default=ROWS_PER_GUIDE,
                    help="trajectory rows per guide per GPU (x %d guides = rows per GPU)" % len(GUIDES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-out", default=None, help="write the per-kernel time table of one UNet forward here")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

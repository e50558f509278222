#!/usr/bin/env python
"""bench.py -- guided-diffusion trajectory sampling throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (CPU arm)
    python bench.py --impl torch-cuda                        # secondary baseline: the same algorithm, stock PyTorch on cuda
    python bench.py --config c3                              # BASELINE configs[2]: 1xB200, guides [1,2,3] x 341 rows

A "step" is ONE pass of the hot path over one batch: the full 255-step guided reverse diffusion
(TemporalUNet + posterior + guide gradient/update) of every trajectory row of the batch, followed
by the per-row best-of-ensemble cost.  Workload (BASELINE.json configs[1]): the
"batch=8192 trajectories, guides [1..10]" ensemble fits one GPU, so at N = 1 it runs whole: the first
ten shipped guides [1,2,3,4,5,9,10,11,12,13] x 819 rows = 8190 trajectories per GPU, 20 synthetic
obstacles, seeded random weights.  Weak scaling (the headline line): every rank runs its own ensemble of
that size (ensembles are rank-local), one NCCL all-gather of the per-row final costs.  The same JSON line
carries a `strong` block: the literal "batch=8192 sharded" reading of configs[1] -- 8190 rows in total,
8190 // N per GPU (1020 rows/GPU at N = 8) -- so that the driver's scaling run records both curves.

Prints ONE JSON line (rank 0).  `value` = trajectories/s with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host x_T in, trajectories + costs out);
`e2e_api` = through the reference-shaped Python call Diffusion.denoise_guided (device noise, and the reference's numpy
noise stream); `gpu_baseline` = stock PyTorch-CUDA eager running the oracle port on the same GPU; `cpu_baseline` = the
oracle port on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1] (headline): guides [1..10] -> the first ten shipped guides
    "c2": {"guides": [1, 2, 3, 4, 5, 9, 10, 11, 12, 13], "rows_per_guide": 819, "name": "configs[1]"},
    # BASELINE.json configs[2]: 1xB200, batch=1024, guides [1,2,3] -> 3 x 341 = 1023 rows
    "c3": {"guides": [1, 2, 3], "rows_per_guide": 341, "name": "configs[2]"},
}
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


def build_workload(guides, rows_per_guide, seed_offset=0):
    """Inputs of one rank: guide tables, scene, x_T, start, goal.  Generated by the product's own synthetic-problem
    module (the CPU arm's oracle has an identical generator; tests/test_host_and_abi.py pins the two together)."""
    from edmp_b200 import build_guide_cfgs, load_guide_hparams, synthetic
    hp = load_guide_hparams(guides, os.path.join(ROOT, "guides") + "/")
    cfgs = build_guide_cfgs(hp, rows_per_guide)
    scene = synthetic.synthetic_scene(N_OBSTACLES, seed=1, rotated=True, cylinders=4)
    rows = cfgs["total_batch_size"]
    # x_T near the straight joint-space line keeps the untrained net's chain finite
    x_T = synthetic.gentle_x_T(rows, synthetic.alpha_bar_T(), seed=100 + seed_offset)
    return cfgs, scene, x_T, synthetic.START.copy(), synthetic.GOAL.copy()


def synthetic_state_dict():
    from edmp_b200 import synthetic                           # seeded random weights (real ones are a download)
    return synthetic.seeded_state_dict(0, final_gain=0.2)


def workload_config(args, n_gpus, precision, rows_per_guide=None, note=None):
    c = CONFIGS[args.config]
    guides = c["guides"]
    rpg = rows_per_guide or args.rows_per_guide
    rows = len(guides) * rpg
    cfg = {"workload": "%s: guided ensemble, guides %s x %d rows = %d rows/GPU (%d total), %d obstacles, "
                       "T=255, horizon 50, 7 DoF" % (c["name"], guides, rpg, rows, rows * n_gpus, N_OBSTACLES),
           "rows_per_gpu": rows, "n_guides": len(guides), "rows_per_guide": rpg,
           "obstacles": N_OBSTACLES, "precision_mode": precision, "parallelism": "dp%d (ensembles rank-local)" % n_gpus,
           "l2": "inputs larger than L2: one UNet forward streams ~57 MB of weight tiles and ~0.25 MB of "
                 "activations per row (%.1f GB at this batch) through the 126 MB L2; no explicit flush" % (rows * 0.25e-3)}
    if note:
        cfg["note"] = note
    return cfg


# -----------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port; /root/reference does not exist on the GPU box)
# -----------------------------------------------------------------------------------------------------
PORT_NOTE = ("CPU arm = the oracle port (torch CPU fp32 UNet + autograd guide, pinned to the reference by the golden "
             "fixtures); per core it is ~5x faster than the survey's probe of the shimmed reference itself (0.44 traj/s on "
             "8 threads at 30 rows, SURVEY.md section 6: the reference builds and discards an autograd graph of the UNet "
             "every step and materialises repeat()ed tensors in the guide), so ratios against it are conservative")


def cpu_sample(guides, rows_per_guide, sd, seed=0):
    from oracle import sampler_oracle as so, guide_oracle as go
    cfgs, scene, x_T, start, goal = build_workload(guides, rows_per_guide, seed)
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
    guides = CONFIGS[args.config]["guides"]
    sd = synthetic_state_dict()
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    cores = torch.get_num_threads()
    # Bounded sample: every timed step is a full 255-step pass over `rpg` rows per guide.  The rows per step are sized
    # from a calibration pass so that the K timed steps end within ~4 minutes: 12 rows per guide (120 rows, SURVEY.md
    # section 8d's largest CPU size) when K is small, fewer when the driver asks for many steps.
    rows_w, dt_w = cpu_sample(guides, 1, sd)                  # warm-up / calibration pass (10 rows)
    for _ in range(max(0, args.warmup - 1)):
        cpu_sample(guides, 1, sd)
    rate = rows_w / dt_w
    rpg = int(max(2, min(12, (240.0 * rate * 1.5) / (max(1, args.steps) * len(guides)))))
    times, rows = [], 0
    for k in range(args.steps):
        rows, dt = cpu_sample(guides, rpg, sd, seed=k)
        times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = rows / (ms / 1000.0)
    sample = "%d rows (%d guides x %d) x 255 steps per timed step, %d obstacles; oracle port (torch CPU fp32 UNet + " \
             "autograd guide), %d threads" % (rows, len(guides), rpg, N_OBSTACLES, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus, "fp32",
                                      note="CPU arm runs a bounded sample of the same workload. " + PORT_NOTE),
            "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# -----------------------------------------------------------------------------------------------------
# secondary baseline: the same algorithm with stock PyTorch kernels on the same GPU (cuDNN / ATen eager)
# -----------------------------------------------------------------------------------------------------
def torch_cuda_pass(sd_dev, cfgs, scene, x_T, start, goal, dev, seed=0, t_start=255, t_stop=0):
    """Steps t_start .. t_stop+1 (default: one whole 255-step pass) of the oracle port (oracle/unet_oracle.py + the autograd guide of oracle/guide_oracle.py) with every
    tensor on `dev`: what a straightforward PyTorch-CUDA port of the reference does per step (~300 library kernels for
    the UNet, autograd for the guide), minus the reference's host round trips and CPU-resident FK temporaries."""
    import torch
    from oracle import guide_oracle as go, sampler_oracle as so, unet_oracle
    rows = cfgs["total_batch_size"]
    beta, alpha, abar = so.schedule()
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.tensor(x_T, dtype=torch.float64, device=dev)
    s_t = torch.tensor(start, dtype=torch.float64, device=dev)
    g_t = torch.tensor(goal, dtype=torch.float64, device=dev)
    X[:, :, 0], X[:, :, -1] = s_t, g_t
    lo = torch.tensor(so.JOINT_LOWER_DEG * (np.pi / 180), device=dev)[None, :, None]
    hi = torch.tensor(so.JOINT_UPPER_DEG * (np.pi / 180), device=dev)[None, :, None]
    sched = torch.tensor(cfgs["guidance_schedule"], device=dev)
    m = torch.tensor(cfgs["guidance_method"], dtype=torch.float32, device=dev).view(rows, 1, 1)
    gn = torch.tensor(cfgs["grad_norm"], dtype=torch.float64, device=dev).view(rows, 1, 1)
    for t in range(t_start, t_stop, -1):
        with torch.no_grad():
            eps = unet_oracle.unet_forward(sd_dev, X.float(), t).double()
            z = torch.randn(X.shape, generator=g, device=dev, dtype=torch.float64)
            a, ab, b = alpha[t - 1], abar[t - 1], beta[t - 1]
            X = (X - ((1 - a) / np.sqrt(1 - ab)) * eps) / np.sqrt(a) + b * z
        if t % 2 == 0 and t >= 5:
            q = torch.clamp(X[:, :, 1:-1], lo, hi).float().requires_grad_(True)
            omin, omax = go.obstacle_aabbs(scene, cfgs["expansion"][:, t - 1], cfgs["clearance"][:, t - 1], rows=rows,
                                           device=dev)
            cost = torch.sum((1 - m) * go.iv_cost(q, omin, omax)) + torch.sum(m * go.sv_cost(q, start, goal, omin, omax))
            cost.backward()
            with torch.no_grad():
                G = q.grad
                mixed = (1 - gn) * G.double() + gn * (G / torch.linalg.norm(G)).double()
                X[:, :, 1:-1] -= sched[:, t - 1, None, None] * mixed
        with torch.no_grad():
            X[:, :, 0], X[:, :, -1] = s_t, g_t
    with torch.no_grad():
        omin, omax = go.obstacle_aabbs(scene, rows=rows, device=dev)
        cost = go.sv_cost(torch.nan_to_num(X[:, :, 1:-1]).float(), start, goal, omin, omax).sum(dim=(1, 2))
    return X, cost


def gpu_baseline_sample(guides, rows_per_guide, dev, sample_steps=8):
    """trajectories/s of the stock PyTorch-CUDA eager port on `dev`, TF32 off (fp32 like the reference) and on.
    A whole pass of it takes minutes at 8190 rows (cuDNN fp32 UNet forward ~190 ms, autograd guide ~1.7 s per guided
    step), so the sample is `sample_steps` consecutive reverse steps of the real loop from t = 254 (half of them guided,
    like 125 of 255 are), extrapolated to 255 steps; the UNet forward alone is timed as well."""
    import torch
    from oracle import unet_oracle
    sd = {k: v.to(dev) for k, v in synthetic_state_dict().items()}
    cfgs, scene, x_T, start, goal = build_workload(guides, rows_per_guide)
    rows = cfgs["total_batch_size"]
    out = {"unit": "trajectories/s", "kind": "torch-cuda eager (oracle port, stock cuDNN / ATen kernels, autograd guide)",
           "rows": rows, "torch": torch.__version__,
           "sample": "%d consecutive reverse steps (t = 254 ..) of the real loop at %d rows, extrapolated x 255 / %d"
                     % (sample_steps, rows, sample_steps)}
    xf = torch.randn(rows, 7, 50, device=dev)
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch_cuda_pass(sd, cfgs, scene, x_T, start, goal, dev, t_start=254, t_stop=252)    # warm-up (cuDNN heuristics)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch_cuda_pass(sd, cfgs, scene, x_T, start, goal, dev, t_start=254, t_stop=254 - sample_steps)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 255.0 / sample_steps
        out[name] = rows / dt
        with torch.no_grad():
            unet_oracle.unet_forward(sd, xf, 100)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                unet_oracle.unet_forward(sd, xf, 100)
            torch.cuda.synchronize()
        out["unet_forward_ms_" + name] = (time.perf_counter() - t0) / 5 * 1e3
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    out["value"] = out["fp32"]
    return out


def run_torch_cuda_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl torch-cuda needs a CUDA device")
    c = CONFIGS[args.config]
    gb = gpu_baseline_sample(c["guides"], args.rows_per_guide, "cuda:0", sample_steps=32)
    line = {"impl": "torch-cuda", "metric": METRIC, "value": gb["value"], "unit": "trajectories/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1, "fp32 (cuDNN / ATen)"),
            "gpu_baseline": gb}
    print(json.dumps(line))


# -----------------------------------------------------------------------------------------------------
# GPU arm
# -----------------------------------------------------------------------------------------------------
class Rank:
    """One rank's engines and inputs for a given rows-per-guide."""

    def __init__(self, args, guides, rows_per_guide, dev, rank, world, sd):
        import torch
        from edmp_b200 import Diffusion, IntersectionVolumeGuide, TemporalUNet
        self.torch = torch
        self.dev, self.rank, self.world = dev, rank, world
        self.model = TemporalUNet(os.path.join(tempfile.mkdtemp(), "TemporalUNetModel255_N50"), 7, 32, dev,
                                  dims=(32, 64, 128, 256, 512, 512), precision=args.precision)
        self.model.load_state_dict(sd)
        self.cfgs, self.scene, self.x_T, self.start, self.goal = build_workload(guides, rows_per_guide, seed_offset=rank)
        self.rows = self.cfgs["total_batch_size"]
        self.guide = IntersectionVolumeGuide(self.scene, dev, self.cfgs, self.rows)
        self.diff = Diffusion(255, dev)
        self.x0 = torch.tensor(self.x_T, dtype=torch.float64, device=dev)
        self.x = torch.empty_like(self.x0)

    def one_pass(self, k):
        """device-resident pass + the single collective (per-row final costs) + best row per ensemble"""
        from edmp_b200 import ensemble
        self.x.copy_(self.x0)
        # (a different Philox stream per pass AND per rank: the counter is the rank-local element index)
        cost = self.diff.run_steps(self.model, self.guide, self.x, self.start, self.goal, 255, 0, noise=None,
                                   seed=k * self.world + self.rank,
                                   guidance_schedule=self.cfgs["guidance_schedule"], want_cost=True)
        allc = ensemble.gather_costs(cost)
        return ensemble.best_rows(allc, self.rows)

    def time_passes(self, steps, warmup, barrier):
        torch = self.torch
        for w in range(warmup):
            self.one_pass(1000 + w)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for k in range(steps):
            self.one_pass(k)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    def time_e2e(self, steps, barrier):
        """through the host-buffer C-ABI call: pinned x_T in, trajectories + costs out, all-gather of the costs"""
        torch = self.torch
        from edmp_b200 import _lib, ensemble
        lib = _lib.load()
        rows = self.rows
        xh = torch.tensor(self.x_T, dtype=torch.float64).pin_memory()
        xh_work = torch.empty_like(xh).pin_memory()
        ch = torch.empty(rows, dtype=torch.float32).pin_memory()
        s_arr, s_ptr = _lib.host_f64(self.start)
        g_arr, g_ptr = _lib.host_f64(self.goal)

        def e2e_pass(seed):
            xh_work.copy_(xh)
            with torch.cuda.device(self.dev):
                _lib.check(lib.edmp_sample_guided_host(
                    self.diff._sampler(rows), self.model.engine(rows),
                    self.guide.scene_handle(rows=rows, guidance_schedule=self.cfgs["guidance_schedule"]),
                    ctypes.c_void_p(xh_work.data_ptr()), s_ptr, g_ptr, ctypes.c_uint64(seed), rows,
                    ctypes.c_void_p(ch.data_ptr()), _lib.stream_ptr()), "edmp_sample_guided_host")
            allc = ensemble.gather_costs(ch.to(self.dev) if self.world > 1 else ch)
            return ensemble.best_rows(allc, rows)

        e2e_pass(5)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            e2e_pass(k * self.world + self.rank)
        barrier()
        return time.perf_counter() - t0


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from edmp_b200 import _lib

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
    c = CONFIGS[args.config]
    guides = c["guides"]
    sd = synthetic_state_dict()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- headline: weak scaling, the whole ensemble per GPU --------------------------------------------------------
    R = Rank(args, guides, args.rows_per_guide, dev, rank, world, sd)
    rows = R.rows
    clocks = ClockSampler(local)
    for w in range(args.warmup):
        R.one_pass(1000 + w)
    barrier()
    if rank == 0:
        clocks.start()
    elapsed_ms = max_over_ranks(R.time_passes(args.steps, 0, barrier))
    launches = R.diff.last_launches * args.steps
    ms_per_step = elapsed_ms / args.steps
    value = rows * world / (ms_per_step / 1000.0)
    e2e_s = max_over_ranks(R.time_e2e(args.steps, barrier))
    e2e_value = rows * world * args.steps / e2e_s
    clk = clocks.stop() if rank == 0 else None
    R.model.check_range()                     # the IEEE-half operand range held through every pass (raises otherwise)

    # ---- strong scaling block: the same 8190-row batch SHARDED over the ranks (configs[1] read literally) ----------
    strong = None
    if not args.no_strong and args.config == "c2":
        rpg_s = max(1, CONFIGS["c2"]["rows_per_guide"] // world)
        if world == 1 and rpg_s == args.rows_per_guide:
            strong = {"value": value, "e2e": e2e_value, "ms_per_step": ms_per_step, "rows_per_gpu": rows,
                      "rows_total": rows, "note": "N = 1: the strong and the weak workload coincide"}
        else:
            S = Rank(args, guides, rpg_s, dev, rank, world, sd)
            ms_s = max_over_ranks(S.time_passes(args.steps, args.warmup, barrier)) / args.steps
            e2e_ss = max_over_ranks(S.time_e2e(args.steps, barrier))
            strong = {"value": S.rows * world / (ms_s / 1000.0), "e2e": S.rows * world * args.steps / e2e_ss,
                      "ms_per_step": ms_s, "rows_per_gpu": S.rows, "rows_total": S.rows * world,
                      "note": "8190 rows in total, sharded: %d guides x %d rows per GPU; total work fixed as N grows"
                              % (len(guides), rpg_s)}
            S.model.check_range()
            del S

    # ---- end to end through the reference-shaped Python API (rank 0, N = 1) ------------------------------------------
    e2e_api = None
    if rank == 0 and world == 1 and not args.no_api_e2e:
        e2e_api = {"unit": "trajectories/s", "rows": rows,
                   "call": "Diffusion.denoise_guided(model, guide, 50, 7, guidance_schedule, batch_size, start, goal)"}
        for mode, passes in (("philox", 2), ("numpy", 1)):
            R.diff.noise_mode = mode
            np.random.seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(passes):
                out = R.diff.denoise_guided(R.model, R.guide, 50, 7, R.cfgs["guidance_schedule"], batch_size=rows,
                                            start=R.start, goal=R.goal, condition=True, benchmarking=True)
                R.guide.choose_best_trajectory(R.start, R.goal, out)
            e2e_api[mode] = rows * passes / (time.perf_counter() - t0)
        e2e_api["note"] = ("philox: x_T drawn on the host, per-step noise on the device; numpy: x_T and all 255 z_t drawn "
                           "by np.random.multivariate_normal in the reference's order (the reference's own stream) and "
                           "streamed to the device in 16-step chunks -- host RNG bound")

    # ---- roofline of the dominant kernel: per-op CUDA-event times of UNet forwards -----------------
    roofline = None
    unet_summary = None
    if rank == 0:
        engine = R.model.engine(rows)
        n_ops = lib.edmp_unet_launches_per_forward(engine)
        ms = np.zeros(n_ops, dtype=np.float32)
        macs = np.zeros(n_ops, dtype=np.float64)
        xf = torch.randn(rows, 7, 50, device=dev)
        eps = torch.empty_like(xf)
        _lib.check(lib.edmp_unet_profile(engine, ctypes.c_void_p(xf.data_ptr()), 128, rows, 10,
                                         ms.ctypes.data_as(ctypes.c_void_p), macs.ctypes.data_as(ctypes.c_void_p),
                                         ctypes.c_void_p(eps.data_ptr()), _lib.stream_ptr()), "edmp_unet_profile")
        names = [lib.edmp_unet_op_name(engine, i).decode() for i in range(n_ops)]
        kernels = [lib.edmp_unet_op_kernel(engine, i).decode() for i in range(n_ops)]
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
        split = prec in ("f16x3", "bf16x3", "tf32x3")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16 sustained%s (%s)" %
                                   (" x 0.5 (tf32-class operands)" if peak != peaks["bf16_tflops_sustained"] else "",
                                    peaks["source"]),
                    "launches_per_forward": n_l, "kernel_ms": float(ms[sel].sum() / n_l),
                    "kernel_share_of_unet": float(ms[sel].sum() / ms.sum()),
                    "useful_flops_per_launch": float(2.0 * macs[sel].sum() / n_l),
                    "frac_of_issued_mma": (3.0 if split else 1.0) * achieved / peak,
                    "note": "useful (non-padding, un-split) FLOPs; the parity-grade hi/lo operand split issues 3x as many "
                            "MMA FLOPs, so `frac` cannot exceed 1/3 in the x3 modes (frac_of_issued_mma counts them)",
                    "slowest_launch": {"op": names[top], "ms": float(ms[top]),
                                       "achieved": float(2.0 * macs[top] / (ms[top] * 1e-3) / 1e12)},
                    "by_kernel_ms": {k: round(v, 4) for k, v in share.items()}}
        if args.ops_out:
            with open(args.ops_out, "w") as f:
                for nm, kn, m, mc in zip(names, kernels, ms, macs):
                    f.write("%-44s %-14s %9.1f us %8.2f useful TFLOP/s\n" % (nm, kn, m * 1e3, 2 * mc / (m * 1e-3) / 1e12 if m > 0 else 0))
        unet_summary = {"ms_per_forward": float(ms.sum()), "useful_tflops": float(2.0 * macs.sum() / (ms.sum() * 1e-3) / 1e12),
                        "launches": int(n_ops)}

    # ---- baselines on this box (rank 0, N = 1 only): stock PyTorch on the same GPU, the oracle port on the host cores ---
    gpu_baseline = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        del R
        torch.cuda.empty_cache()
        try:
            gpu_baseline = gpu_baseline_sample(guides, args.rows_per_guide, dev)
        except Exception as e:   # noqa: BLE001  (a baseline must not take the bench line down, e.g. out of memory)
            gpu_baseline = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = torch.get_num_threads()
        rpg_c = min(12, args.rows_per_guide)
        crow, cdt = cpu_sample(guides, rpg_c, sd)
        cpu_baseline = {"value": crow / cdt, "unit": "trajectories/s", "cores": cores, "kind": "port",
                        "sample": "%d rows (%d guides x %d) x 255 steps, %d obstacles, %.1f s; oracle port "
                                  "(torch CPU fp32 UNet + autograd guide)" % (crow, len(guides), rpg_c, N_OBSTACLES, cdt),
                        "note": PORT_NOTE}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else args.precision, "data": "synthetic",
                "config": workload_config(args, world, args.precision),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": rows * 350 * 8,
                        "d2h_bytes_per_step": rows * 350 * 8 + rows * 4},
                "gpu_launches": int(launches),
                "strong": strong, "e2e_api": e2e_api,
                "roofline": roofline, "unet": unet_summary, "gpu_baseline": gpu_baseline, "cpu_baseline": cpu_baseline,
                "useful_tflops_whole_job": value * 255 * USEFUL_GFLOP_PER_ROW_STEP / 1e3}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-cuda"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="c2 = BASELINE configs[1] (headline, default); c3 = configs[2] (1xB200, guides [1,2,3] x 341)")
    ap.add_argument("--precision", default=os.environ.get("EDMP_PRECISION", "f16x3"),
                    help="f16x3 (tcgen05, IEEE-half hi/lo split, parity grade, default) | tf32x3 | fp32 (CUDA cores) | "
                         "bf16x3 / f16 / bf16 / tf32 (not parity grade)")
    ap.add_argument("--rows-per-guide", type=int, default=None,
                    help="trajectory rows per guide per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block")
    ap.add_argument("--no-api-e2e", action="store_true", help="skip the Python-API end-to-end timings")
    ap.add_argument("--quick", action="store_true", help="development runs: only the headline numbers and the roofline")
    ap.add_argument("--ops-out", default=None, help="write the per-kernel time table of one UNet forward here")
    args = ap.parse_args()
    if args.rows_per_guide is None:
        args.rows_per_guide = CONFIGS[args.config]["rows_per_guide"]
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_strong = args.no_api_e2e = True
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "torch-cuda":
        run_torch_cuda_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

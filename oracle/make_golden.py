"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, under oracle/ref_shim.py) on seeded inputs.  Run in the build container:

    python -m oracle.make_golden

The fixtures pin (a) the oracle restatement and (b) the CUDA path; the GPU box has no
/root/reference, so only these files travel.  Everything is regenerated deterministically
from seeds recorded inside each file (weights: oracle/weights.py; scenes/noise: numpy
default_rng).
"""
import os
import sys

import numpy as np
import torch

from . import guide_params, ref_shim, sampler_oracle as so, scenes, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
UNET_TAPS = ["down_samplers.0.down.0", "down_samplers.1.down.1", "down_samplers.2.down.3",
             "down_samplers.5.down.1", "middle_block.middle.2", "up_samplers.0.up.3",
             "up_samplers.1.up.3", "up_samplers.2.up.0", "up_samplers.4.up.3", "final_conv.0"]


def _module_by_path(model, path):
    m = model
    for part in path.split("."):
        m = m[int(part)] if part.isdigit() else getattr(m, part)
    return m


def golden_unet():
    sd = weights.seeded_state_dict(0)
    model = ref_shim.make_reference_unet(sd)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 7, 50, generator=g)
    out = {"weights_seed": 0, "x": x.numpy()}
    taps = {}
    hooks = []
    for p in UNET_TAPS:
        def hook(mod, inp, res, p=p):
            taps[p] = (res[0] if isinstance(res, tuple) else res).detach().numpy().copy()
        hooks.append(_module_by_path(model, p).register_forward_hook(hook))
    for t in (255, 128, 1):
        with torch.no_grad():
            out["eps_t%d" % t] = model(x, torch.tensor([float(t)])).numpy()
        if t == 128:
            for p, v in taps.items():
                out["tap_t128/" + p] = v
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(OUT, "unet_forward.npz"), **out)
    print("unet_forward.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def _noisy_line(B, rng, sigma=0.3):
    line = scenes.START[None, :, None] + (scenes.GOAL - scenes.START)[None, :, None] * \
        np.linspace(0, 1, 50)[None, None, 1:-1]
    return so.clip_joints(line + sigma * rng.normal(size=(B, 7, 48)))


def golden_guide():
    ns = ref_shim.load_reference()
    out = {"link_dims": ref_shim.reference_link_dimensions()}
    cases = {"mixed": ([1, 10, 11, 9], 3, scenes.synthetic_scene(8, seed=1, rotated=True, cylinders=2)),
             "iv": ([1, 2, 3], 2, scenes.synthetic_scene(20, seed=2, rotated=True)),
             "sv_axis": ([10, 13, 18], 2, scenes.synthetic_scene(5, seed=4, rotated=False))}
    for name, (guides, bpg, scene) in cases.items():
        cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
        B = cfgs["total_batch_size"]
        rng = np.random.default_rng(100 + len(name))
        q = _noisy_line(B, rng)
        guide = ns.IntersectionVolumeGuide(obstacle_config=scene, device="cpu", guide_cfgs=cfgs,
                                           batch_size=B)
        out[name + "/guides"] = np.array(guides)
        out[name + "/bpg"] = bpg
        out[name + "/scene"] = scene
        out[name + "/q"] = q
        for t in (254, 100, 6):
            out["%s/grad_t%d" % (name, t)] = guide.get_gradient(q, scenes.START, scenes.GOAL, t)
        # t = 0 costs: goal filter (infer_serial.py:119) and best-of-ensemble (lib/guide.py:637)
        ik = q[:, :, :1]
        out[name + "/cost_t0"] = guide.cost(torch.tensor(ik), 0, batch_size=B).numpy()
        qt = torch.tensor(q, dtype=torch.float32)
        out[name + "/iv_t100"] = guide.cost(qt, 100).numpy()
        out[name + "/sv_t100"] = guide.swept_volume_cost(
            qt, torch.tensor(scenes.START, dtype=torch.float32),
            torch.tensor(scenes.GOAL, dtype=torch.float32), 100).numpy()
        traj = np.concatenate([np.broadcast_to(scenes.START[None, :, None], (B, 7, 1)), q,
                               np.broadcast_to(scenes.GOAL[None, :, None], (B, 7, 1))], axis=2)
        best = guide.choose_best_trajectory(scenes.START, scenes.GOAL, traj)
        out[name + "/best_index"] = int(np.argmin(np.abs(traj - best[None]).sum(axis=(1, 2))))
        out[name + "/final_sv"] = torch.sum(guide.swept_volume_cost(
            torch.tensor(traj[:, :, 1:-1], dtype=torch.float32),
            torch.tensor(scenes.START, dtype=torch.float32),
            torch.tensor(scenes.GOAL, dtype=torch.float32), 0), dim=(1, 2)).numpy()
    # NaN poisoning: a grad_norm row in a batch whose gradient is identically zero (:629)
    far = np.array([[5.0, 5.0, 5.0, 0, 0, 0, 1, 0.1, 0.1, 0.1]])
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in (1, 9)], 1)
    guide = ns.IntersectionVolumeGuide(obstacle_config=far, device="cpu", guide_cfgs=cfgs, batch_size=2)
    q = _noisy_line(2, np.random.default_rng(9))
    out["nan/scene"], out["nan/q"] = far, q
    with np.errstate(all="ignore"):
        out["nan/grad_t100"] = guide.get_gradient(q, scenes.START, scenes.GOAL, 100)
    np.savez_compressed(os.path.join(OUT, "guide.npz"), **out)
    print("guide.npz", len(out), "arrays; nan case all-nan:", np.isnan(out["nan/grad_t100"]).all())


def _run_reference_sampler(guides, bpg, scene, sd, x_T, noise, record_steps):
    ns = ref_shim.load_reference()
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    B = cfgs["total_batch_size"]
    model = ref_shim.make_reference_unet(sd)
    guide = ns.IntersectionVolumeGuide(obstacle_config=scene, device="cpu", guide_cfgs=cfgs, batch_size=B)
    diff = ns.Diffusion(T=255, device="cpu")
    ref_shim.set_noise_tape(ref_shim.NoiseTape(replay=[x_T] + list(noise)))
    rec = {}
    post, grad = diff.p_sample_using_posterior, guide.get_gradient

    def post_wrap(xt, t, eps):
        res = post(xt, t, eps)
        if t in record_steps:
            rec["x_in_t%d" % t], rec["eps_t%d" % t], rec["x_post_t%d" % t] = xt.copy(), eps.copy(), res.copy()
        if (t + 1) in record_steps:
            rec["x_out_t%d" % (t + 1)] = xt.copy()
        return res

    def grad_wrap(joint_input, start, goal, t):
        res = grad(joint_input, start, goal, t)
        if t in record_steps:
            rec["grad_t%d" % t] = res.copy()
        return res

    diff.p_sample_using_posterior = post_wrap
    guide.get_gradient = grad_wrap
    devnull = open(os.devnull, "w")
    stdout, sys.stdout = sys.stdout, devnull
    try:
        final = diff.denoise_guided(model=model, guide=guide, batch_size=B, traj_len=50, num_channels=7,
                                    condition=True, benchmarking=True, start=scenes.START,
                                    goal=scenes.GOAL, guidance_schedule=cfgs["guidance_schedule"])
    finally:
        sys.stdout = stdout
    if 1 in record_steps:
        rec["x_out_t1"] = final.copy()
    return final, rec, guide


def golden_sampler():
    beta, alpha, abar = so.schedule()
    base = scenes.tabletop_scene()
    scene = np.vstack([base, np.array([[0.12, 0, 0.2, 0, 0, 0, 1, 0.1, 0.1, 0.1]])])
    sd = weights.seeded_state_dict(0, final_gain=0.2)
    steps = (255, 254, 253, 200, 151, 150, 100, 81, 80, 21, 20, 6, 5, 2, 1)
    for name, guides, bpg, xseed, nseed in (("iv", [1, 2, 3], 2, 5, 7), ("mixed", [1, 10, 11, 9], 1, 6, 8)):
        B = len(guides) * bpg
        x_T = scenes.gentle_x_T(B, abar[-1], seed=xseed)
        rng = np.random.default_rng(nseed)
        noise = [rng.normal(size=(B, 7, 50)) for _ in range(255)]
        final, rec, guide = _run_reference_sampler(guides, bpg, scene, sd, x_T, noise, steps)
        best = guide.choose_best_trajectory(scenes.START, scenes.GOAL, final)
        out = {"guides": np.array(guides), "bpg": bpg, "scene": scene, "weights_seed": 0, "final_gain": 0.2,
               "x_T_seed": xseed, "noise_seed": nseed, "x_T": x_T, "final": final,
               "noise_checksum": float(sum(n.sum() for n in noise)), "best": best,
               "steps": np.array(steps)}
        out.update(rec)
        np.savez_compressed(os.path.join(OUT, "sampler_%s.npz" % name), **out)
        print("sampler_%s.npz" % name, "final range", final.min(), final.max(),
              "nan" if np.isnan(final).any() else "finite")


# ---- every one of the 255 steps, teacher-forced -----------------------------------------------------------------
def _run_reference_tape(guides, bpg, scene, sd, x_T, noise, condition=True):
    """The reference's denoise_guided with ONE intervention by the harness: the state is rounded to float32 at the entry
    of every p_sample_using_posterior call (the network input is float32 anyway, diffusion.py:319).  Every single step
    is then the UNMODIFIED reference step function applied to a float32-representable state, so ONE float32 tape
    [256, B, 7, 50] holds the exact input AND (to 1e-7 relative) the reference's output of all 255 steps:
    tape[k] = state entering step t = 255 - k, tape[k + 1] = reference output of that step."""
    ns = ref_shim.load_reference()
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    B = cfgs["total_batch_size"]
    model = ref_shim.make_reference_unet(sd)
    guide = ns.IntersectionVolumeGuide(obstacle_config=scene, device="cpu", guide_cfgs=cfgs, batch_size=B)
    diff = ns.Diffusion(T=255, device="cpu")
    ref_shim.set_noise_tape(ref_shim.NoiseTape(replay=[x_T] + list(noise)))
    tape = np.zeros((256, B, 7, 50), dtype=np.float32)
    post = diff.p_sample_using_posterior

    def post_wrap(xt, t, eps):
        x32 = xt.astype(np.float32)
        tape[255 - t] = x32
        return post(x32.astype(np.float64), t, eps)

    diff.p_sample_using_posterior = post_wrap
    devnull = open(os.devnull, "w")
    stdout, sys.stdout = sys.stdout, devnull
    try:
        with np.errstate(all="ignore"):
            final = diff.denoise_guided(model=model, guide=guide, batch_size=B, traj_len=50, num_channels=7,
                                        condition=condition, benchmarking=True, start=scenes.START,
                                        goal=scenes.GOAL, guidance_schedule=cfgs["guidance_schedule"])
    finally:
        sys.stdout = stdout
    tape[255] = final.astype(np.float32)
    return tape, final


def _oracle_divergence(guides, bpg, scene, sd, x_T, noise, final_ref):
    """per-row max |oracle chain - reference chain| at t = 0: how far the reference's OWN arithmetic, re-ordered
    (torch fp32 restatement + autograd guide), drifts over 255 free-running steps.  The bound the GPU chain is held to
    on the chaotic (sv / grad-norm) rows is tied to it."""
    cfgs = so.expand_guide_tables([guide_params.GUIDES[n] for n in guides], bpg)
    with np.errstate(all="ignore"):
        out = so.denoise_guided(sd, scene, cfgs, scenes.START, scenes.GOAL, x_T, noise, gradient="autograd")
    return np.abs(out - final_ref).max(axis=(1, 2))


TAPE_CASES = {
    # name: (guides, rows per guide, scene, final_gain, x_T kind, x_T seed, noise seed, condition)
    "iv": ([1, 2, 3], 2, "tabletop", 0.2, "gentle", 5, 7, True),
    "mixed": ([1, 10, 11, 9], 1, "tabletop", 0.2, "gentle", 6, 8, True),
    # the bench workload's ensemble and scene (bench.py build_workload), one row per guide
    "bench10": ([1, 2, 3, 4, 5, 9, 10, 11, 12, 13], 1, "bench20", 0.2, "gentle", 100, 9, True),
    # the chain as diffusion.py:303 draws it: x_T ~ N(0, I), un-scaled weights
    "normal": ([1, 10, 11, 9], 1, "tabletop", 1.0, "normal", 12, 13, True),
    # condition=False (diffusion.py:300-301,:347): the endpoints diffuse freely
    "uncond": ([1, 10], 1, "tabletop", 0.2, "gentle", 14, 15, False),
}


def tape_case_inputs(name):
    """(guides, bpg, scene, state_dict, x_T, noise list, condition) of a tape case, regenerated from its seeds."""
    guides, bpg, scene_kind, final_gain, x_kind, xseed, nseed, condition = TAPE_CASES[name]
    beta, alpha, abar = so.schedule()
    if scene_kind == "tabletop":
        scene = np.vstack([scenes.tabletop_scene(), np.array([[0.12, 0, 0.2, 0, 0, 0, 1, 0.1, 0.1, 0.1]])])
    else:
        scene = scenes.synthetic_scene(20, seed=1, rotated=True, cylinders=4)
    sd = weights.seeded_state_dict(0, final_gain=final_gain)
    B = len(guides) * bpg
    if x_kind == "gentle":
        x_T = scenes.gentle_x_T(B, abar[-1], seed=xseed)
    else:
        x_T = np.random.default_rng(xseed).normal(size=(B, 7, 50))
    rng = np.random.default_rng(nseed)
    noise = [rng.normal(size=(B, 7, 50)) for _ in range(255)]
    return guides, bpg, scene, sd, x_T, noise, condition


def golden_tapes(which=None):
    for name in (which or TAPE_CASES):
        guides, bpg, scene, sd, x_T, noise, condition = tape_case_inputs(name)
        tape, final = _run_reference_tape(guides, bpg, scene, sd, x_T, noise, condition)
        out = {"guides": np.array(guides), "bpg": bpg, "scene": scene, "condition": int(condition),
               "tape": tape, "noise_checksum": float(sum(n.sum() for n in noise))}
        if name in ("mixed", "bench10"):
            # free-running drift of the reference's own arithmetic re-ordered, on the true (un-rounded) chain
            g = np.load(os.path.join(OUT, "sampler_mixed.npz")) if name == "mixed" else None
            final_true = g["final"] if g is not None else _run_reference_sampler(guides, bpg, scene, sd, x_T, noise, ())[0]
            out["final_true"] = final_true
            out["oracle_divergence"] = _oracle_divergence(guides, bpg, scene, sd, x_T, noise, final_true)
        np.savez_compressed(os.path.join(OUT, "tape_%s.npz" % name), **out)
        print("tape_%s.npz" % name, tape.shape, "range", float(np.nanmin(tape)), float(np.nanmax(tape)),
              "nan" if np.isnan(tape).any() else "finite",
              "oracle divergence %s" % np.array2string(out["oracle_divergence"], precision=2) if "oracle_divergence" in out else "")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["unet", "guide", "sampler"]
    if "unet" in which:
        golden_unet()
    if "guide" in which:
        golden_guide()
    if "sampler" in which:
        golden_sampler()
    if "tapes" in which or any(w.startswith("tape:") for w in which):
        golden_tapes([w.split(":", 1)[1] for w in which if w.startswith("tape:")] or None)

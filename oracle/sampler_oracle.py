"""TEST INFRASTRUCTURE ONLY.  numpy-float64 restatement of the guided reverse-diffusion loop.

Follows reference diffusion/diffusion.py: schedule :10-20,:37-49; p_sample_using_posterior
:116-135 (adds beta*z, and at t==1 zeroes the noise of batch row 0 only -- numpy-1.x
semantics of ``z[np.where(t == 1)] = 0``); clip_joints :280-298;
denoise_guided :300-356.  Per-row guide tables follow infer_serial.py:56-91.
"""
import numpy as np
import torch

from . import guide_oracle, unet_oracle

JOINT_LOWER_DEG = np.array([-166.0, -101.0, -166.0, -176.0, -166.0, -1.0, -166.0])
JOINT_UPPER_DEG = np.array([166.0, 101.0, 166.0, -4.0, 166.0, 215.0, 166.0])


def schedule(T=255, thresh=0.02):
    beta = np.linspace(0, thresh, T + 1)[1:]
    alpha = 1 - beta
    alpha_bar = np.array([np.prod(alpha[:t]) for t in range(1, T + 1)])
    return beta, alpha, alpha_bar


def clip_joints(joints):
    lo = JOINT_LOWER_DEG * (np.pi / 180)
    hi = JOINT_UPPER_DEG * (np.pi / 180)
    return np.clip(joints, lo[None, :, None], hi[None, :, None])


def expand_guide_tables(guide_hparams, batch_size_per_guide, T=255):
    """guide_hparams: list of the ``hyperparameters`` dicts of guide<N>.yaml, in ensemble order.
    Returns the guide_cfgs dict infer_serial.py:59-91 builds."""
    G, bpg = len(guide_hparams), batch_size_per_guide
    R = G * bpg
    cfg = {"batch_size_per_guide": bpg, "total_batch_size": R,
           "clearance": np.zeros((R, T)), "expansion": np.zeros((R, T)),
           "guidance_method": np.zeros((R,)), "grad_norm": np.zeros((R,)),
           "guidance_schedule": np.zeros((R, T)), "volume_trust_region": np.zeros((R,))}
    for i, h in enumerate(guide_hparams):
        rows = slice(i * bpg, (i + 1) * bpg)
        rng = h["obstacle_clearance"]["range"]
        cfg["clearance"][rows, :] = np.linspace(rng[0], rng[1], T)
        oe = h["obstacle_expansion"]
        for k in ("1", "2", "3"):                      # later segments overwrite earlier ones
            a, b = oe["isr" + k]
            v = oe["val" + k]
            cfg["expansion"][rows, a:b] = np.linspace(v[0], v[1], num=abs(b - a))
        cfg["guidance_method"][rows] = 1 if h["guidance_method"] == "sv" else 0
        cfg["grad_norm"][rows] = 1 if h["grad_norm"] else 0
        gs = h["guidance_schedule"]
        cfg["guidance_schedule"][rows, :] = (1.4 + np.arange(T) / T) if gs["type"] == "varying" \
            else gs["scale_val"]
        cfg["volume_trust_region"][rows] = h["volume_trust_region"]
    return cfg


def posterior_step(xt, t, eps, z, beta, alpha, alpha_bar):
    z = z.copy()
    if t == 1:
        z[0, :, :] = 0                                 # row 0 only (D6)
    a, ab, b = alpha[t - 1], alpha_bar[t - 1], beta[t - 1]
    return (xt - ((1 - a) / np.sqrt(1 - ab)) * eps) / np.sqrt(a) + b * z


def denoise_guided(state_dict, obstacle_config, guide_cfgs, start, goal, x_T, noise, T=255,
                   gradient="autograd", link_dims=guide_oracle.LINK_DIMS, record=None,
                   model_fn=None):
    """x_T [B,7,50] f64 (before endpoint conditioning), noise[k] = z for step t = T-k.
    ``record``: optional dict receiving per-step tensors for teacher-forced parity:
    record['x_in'][t], ['eps'][t], ['x_post'][t], ['grad'][t], ['x_out'][t]."""
    beta, alpha, alpha_bar = schedule(T)
    grad_fn = guide_oracle.gradient_autograd if gradient == "autograd" else guide_oracle.gradient_analytic
    X = np.array(x_T, dtype=np.float64, copy=True)
    X[:, :, 0] = start
    X[:, :, -1] = goal
    sched = guide_cfgs["guidance_schedule"]
    for step, t in enumerate(range(T, 0, -1)):
        xin = torch.tensor(X, dtype=torch.float32)
        with torch.no_grad():
            eps = (model_fn(xin, t) if model_fn is not None
                   else unet_oracle.unet_forward(state_dict, xin, t)).numpy()
        if record is not None:
            record.setdefault("x_in", {})[t] = X.copy()
            record.setdefault("eps", {})[t] = eps.copy()
        X = posterior_step(X, t, eps, noise[step], beta, alpha, alpha_bar)
        if record is not None:
            record.setdefault("x_post", {})[t] = X.copy()
        if t % 2 == 0 and t >= 5:
            clipped = clip_joints(X[:, :, 1:-1])
            G = grad_fn(clipped, start, goal, obstacle_config, guide_cfgs, t, link_dims)
            X[:, :, 1:-1] = X[:, :, 1:-1] - sched[:, t - 1, None, None] * G
            if record is not None:
                record.setdefault("grad", {})[t] = G.copy()
        X[:, :, 0] = start
        X[:, :, -1] = goal
        if record is not None:
            record.setdefault("x_out", {})[t] = X.copy()
    return X.copy()

"""Debug: one teacher-forced step of a tape case on the GPU, split into UNet / guide parts, against the CPU oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import guide_oracle as go, guide_params, sampler_oracle as so, scenes, unet_oracle, weights
from oracle.make_golden import TAPE_CASES, tape_case_inputs
from edmp_b200 import Diffusion, IntersectionVolumeGuide, TemporalUNet, build_guide_cfgs
import tempfile
case = sys.argv[1]; steps = [int(v) for v in sys.argv[2].split(",")]; prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
g = np.load(os.path.join(ROOT, "tests", "golden", "tape_%s.npz" % case))
guides, bpg, scene, sd, x_T, noise, condition = tape_case_inputs(case)
cfgs = build_guide_cfgs([guide_params.GUIDES[n] for n in guides], bpg)
B = cfgs["total_batch_size"]
m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, "cuda:0", dims=(32, 64, 128, 256, 512, 512), precision=prec)
m.load_state_dict(sd)
guide = IntersectionVolumeGuide(scene, "cuda:0", cfgs, B)
beta, alpha, abar = so.schedule()
tape = g["tape"].astype(np.float64)
np.set_printoptions(precision=2, linewidth=200)
for t in steps:
    k = 255 - t
    X = tape[k]
    xin = torch.tensor(X, dtype=torch.float32)
    with torch.no_grad():
        eps_ref = unet_oracle.unet_forward(sd, xin, t).numpy()
    eps = m(xin.cuda(), t).cpu().numpy()
    print("t", t, "eps err/row", np.abs(eps - eps_ref).max(axis=(1, 2)), "|eps|", np.abs(eps_ref).max(axis=(1, 2)), "|x|", np.abs(X).max(axis=(1, 2)))
    Xp = so.posterior_step(X, t, eps_ref, noise[k], beta, alpha, abar)
    q = so.clip_joints(Xp[:, :, 1:-1])
    if t % 2 == 0 and t >= 5:
        with np.errstate(all="ignore"):
            Ga = go.gradient_analytic(q, scenes.START, scenes.GOAL, scene, cfgs, t)
        Gg = guide.get_gradient(q, scenes.START, scenes.GOAL, t)
        print("   grad err/row", np.abs(Gg - Ga).max(axis=(1, 2)), "|G|", np.abs(Ga).max(axis=(1, 2)))
    d = Diffusion(255, "cuda:0")
    x = torch.tensor(X).cuda()
    d.run_steps(m, guide, x, scenes.START, scenes.GOAL, t, t - 1, noise=torch.tensor(noise[k][None]).cuda(), guidance_schedule=cfgs["guidance_schedule"], condition=condition)
    print("   step err/row vs tape", np.abs(x.cpu().numpy() - tape[k + 1]).max(axis=(1, 2)))

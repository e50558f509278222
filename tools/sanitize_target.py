"""Small target for compute-sanitizer (memcheck): one UNet forward at ragged row counts (single-CTA and, with
EDMP_CG2=1, CTA-pair layers), three guided sampler steps and the sphere-SDF kernels.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import Diffusion, IntersectionVolumeGuide, TemporalUNet, build_guide_cfgs, load_guide_hparams, synthetic  # noqa: E402
from edmp_b200.lib import SphereSDFGuide  # noqa: E402

dev = "cuda:0"
sd = synthetic.seeded_state_dict(0, final_gain=0.2)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 300
m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, dev, dims=(32, 64, 128, 256, 512, 512), precision="f16x3")
m.load_state_dict(sd)
x = torch.randn(rows, 7, 50, device=dev)
eps = m(x, 100)
torch.cuda.synchronize()
print("forward ok", float(eps.abs().max()))
if len(sys.argv) > 2 and sys.argv[2] == "unet":     # large batches: the multi-layer persistent runs, UNet only
    eps2 = m(x, 100)
    torch.cuda.synchronize()
    assert torch.equal(eps, eps2)
    print("runs ok (bit-reproducible)")
    sys.exit(0)
hp = load_guide_hparams([1, 10, 9], os.path.join(ROOT, "guides") + "/")
cfgs = build_guide_cfgs(hp, 2)
guide = IntersectionVolumeGuide(synthetic.tabletop_scene(), dev, cfgs, cfgs["total_batch_size"])
diff = Diffusion(255, dev)
xs = torch.tensor(synthetic.gentle_x_T(6, synthetic.alpha_bar_T(), seed=4), device=dev)
diff.run_steps(m, guide, xs, synthetic.START, synthetic.GOAL, 254, 251, noise=None, seed=1, guidance_schedule=cfgs["guidance_schedule"])
torch.cuda.synchronize()
print("sampler steps ok")
g = SphereSDFGuide(synthetic.synthetic_scene(12, seed=2), None, dev)
q = torch.tensor(synthetic.gentle_x_T(9, 1.0, seed=3, spread=0.2), dtype=torch.float32, device=dev)
c, gr, cl = g.evaluate(q)
cc = g.cloud_clearance(np.random.default_rng(0).uniform(-0.5, 0.9, size=(1500, 3)), q)
torch.cuda.synchronize()
print("sdf ok", float(c.sum()), float(cc.min()))
from edmp_b200.lib import MetricsCalculator  # noqa: E402
mc = MetricsCalculator(guide)
traj = np.cumsum(np.random.default_rng(1).normal(scale=0.03, size=(37, 7, 50)), axis=2)
r = mc.ensemble_metrics(traj, 0.04, return_spectra=True)
s = mc.sparc(np.exp(-5 * np.arange(-1, 1, 0.01) ** 2), fs=100.)
T = guide.get_end_effector_transform(guide.rearrange_joints(torch.tensor(traj, dtype=torch.float32, device=dev)))
torch.cuda.synchronize()
print("metrics ok", float(r["joint_smoothness"].mean()), s[0], tuple(T.shape))

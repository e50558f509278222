"""Per-layer comparison of the CUDA UNet against the torch oracle (all named activations)."""
import os, sys, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import TemporalUNet
from oracle import unet_oracle, weights
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 21
sd = weights.seeded_state_dict(0)
m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, "cuda:0", dims=(32, 64, 128, 256, 512, 512), precision=prec)
m.load_state_dict(sd)
x = torch.randn(rows, 7, 50, generator=torch.Generator().manual_seed(1)) * 1.5
taps = {}
with torch.no_grad():
    ref = unet_oracle.unet_forward(sd, x, 77, taps=taps) if "taps" in unet_oracle.unet_forward.__code__.co_varnames else unet_oracle.unet_forward(sd, x, 77)
eps = m(x.cuda(), 77).cpu()
print("eps max err", (eps - ref).abs().max().item(), "ref max", ref.abs().max().item())
for name, r in taps.items():
    try:
        a = m.read_activation(name, rows).cpu()
    except Exception as e:
        print("%-40s unavailable (%s)" % (name, str(e)[:60])); continue
    r = r[:, :, :a.shape[2]]
    print("%-40s shape %-18s max err %.3e  (ref max %.3f)" % (name, tuple(a.shape), (a - r).abs().max().item(), r.abs().max().item()))

"""Layer-by-layer comparison of the tensor-core engine against the fp32 CUDA-core engine (GPU)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import TemporalUNet  # noqa: E402
from oracle import weights  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
dev = "cuda:0"
sd = weights.seeded_state_dict(0)
models = {}
for p in ("fp32", prec):
    m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, dev, dims=(32, 64, 128, 256, 512, 512), precision=p)
    m.load_state_dict(sd)
    models[p] = m
x = torch.randn(rows, 7, 50, generator=torch.Generator().manual_seed(1)).to(dev)
out = {p: m(x, 128) for p, m in models.items()}
torch.cuda.synchronize()
names = []
for i in range(6):
    for k in (0, 1):
        names += ["down_samplers.%d.down.%d.blocks.0" % (i, k), "down_samplers.%d.down.%d" % (i, k)]
    if i != 5:
        names.append("down_samplers.%d.down.3" % i)
for k in (0, 2):
    names += ["middle_block.middle.%d.blocks.0" % k, "middle_block.middle.%d" % k]
for i in range(5):
    for k in (0, 1):
        names += ["up_samplers.%d.up.%d.blocks.0" % (i, k), "up_samplers.%d.up.%d" % (i, k)]
    names.append("up_samplers.%d.up.3" % i)
names.append("final_conv.0")
for n in names:
    a = models["fp32"].read_activation(n, rows)
    b = models[prec].read_activation(n, rows)
    err = (a - b).abs().max().item()
    print("%-40s %-18s ref absmax %.3f  err %.3g%s" % (n, tuple(a.shape), a.abs().max().item(), err,
                                                     "   <-- BAD" if not err < 1e-3 else ""))
print("eps err", (out["fp32"] - out[prec]).abs().max().item())

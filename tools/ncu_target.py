"""Small target for ncu: a few guided reverse steps at the bench's batch size (1030 rows).
Usage: ncu ... python tools/ncu_target.py [rows] [steps] [precision]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from edmp_b200 import Diffusion, IntersectionVolumeGuide, TemporalUNet  # noqa: E402

rows_per_guide = int(sys.argv[1]) // 10 if len(sys.argv) > 1 else 103
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = "cuda:0"
cfgs, scene, x_T, start, goal = bench.build_workload(bench.CONFIGS["c2"]["guides"], rows_per_guide)
rows = cfgs["total_batch_size"]
model = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, dev, dims=(32, 64, 128, 256, 512, 512),
                     precision=precision)
model.load_state_dict(bench.synthetic_state_dict())
guide = IntersectionVolumeGuide(scene, dev, cfgs, rows)
diff = Diffusion(255, dev)
x = torch.tensor(x_T, device=dev)
diff.run_steps(model, guide, x, start, goal, 200, 200 - steps, noise=None, seed=1,
               guidance_schedule=cfgs["guidance_schedule"])
torch.cuda.synchronize()
print("ran", steps, "steps on", rows, "rows;", diff.last_launches, "launches")

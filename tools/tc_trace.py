"""Per-CTA phase timing (SM clocks) of tensor-core ops: where does a conv_tc_kernel CTA spend time?"""
import ctypes
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import TemporalUNet, _lib  # noqa: E402
from oracle import weights  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1020
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
dev = "cuda:0"
m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, dev, dims=(32, 64, 128, 256, 512, 512), precision=prec)
m.load_state_dict(weights.seeded_state_dict(0))
x = torch.randn(rows, 7, 50, device=dev)
m(x, 100)
torch.cuda.synchronize()
lib = _lib.load()
h = m.engine(rows)
n_ops = lib.edmp_unet_launches_per_forward(h)
buf = np.zeros((8192, 16), dtype=np.int64)
labels = ["setup", "wait W0", "wait A0", "mainloop issue", "acc ready", "epilogue", "teardown"]
for i in range(n_ops - 1):
    n = ctypes.c_int()
    rc = lib.edmp_unet_tc_trace(h, i, rows, buf.ctypes.data_as(ctypes.c_void_p), 8192, ctypes.byref(n), None)
    if rc != 0:
        continue
    t = buf[:n.value].astype(np.float64)
    extra = t[:, 8:12].mean(axis=0)
    t = t[:, :8]
    d = np.diff(t, axis=1)
    total = t[:, 7] - t[:, 0]
    span = (t[:, 7].max() - t[:, 0].min())
    raw = buf[:n.value].astype(np.float64)
    if 0 < raw[:, 5].mean() < raw[:, 0].mean():   # persistent kernels (conv_tc2 / conv_pm2): slots 2..6, 8.. are cycle counters, not clock stamps
        lead = raw[raw[:, 5] > 0]   # CTAs whose MMA warp issued (pair leaders / all CTAs of single-CTA layers)
        r = raw.mean(axis=0)
        r[2:6] = lead[:, 2:6].mean(axis=0)
        print("%-36s ctas %4d  kernel %7.0f cyc (%5.1f us) | setup %5.0f  mma-warp total %7.0f  wait acc_empty %6.0f  wait W %6.0f  wait A %6.0f | epilogue busy %7.0f (stats %6.0f bar %6.0f final %6.0f) params %6.0f  wait acc_full %7.0f"
              % (lib.edmp_unet_op_name(h, i).decode(), n.value, r[7] - r[0], (r[7] - r[0]) / 1900.0, r[1] - r[0], r[5], r[2], r[3], r[4], r[8], r[9], r[10], r[11], r[12], r[6]))
        continue
    print("%-36s ctas %4d  span %8.0f  total %7.0f cyc (%5.1f us @1.9GHz) | " % (lib.edmp_unet_op_name(h, i).decode(), n.value,
          span, total.mean(), total.mean() / 1900.0) + "  ".join("%s %6.0f" % (l, v) for l, v in zip(labels, d.mean(axis=0))) + "  | waitA %6.0f waitB %6.0f a_it %3.0f b_it %3.0f" % tuple(extra))

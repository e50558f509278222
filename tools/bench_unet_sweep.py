"""TemporalUNet255-only forward (no guidance) across batch sizes (BASELINE.json configs[3]): per-forward device time from
CUDA events around every launch (edmp_unet_profile), useful TFLOP/s (61,096,192 non-padding MACs per row) and the
fraction of the measured sustained bf16 peak.   python tools/bench_unet_sweep.py [precision] [out.txt]"""
import ctypes
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edmp_b200 import TemporalUNet, _lib, synthetic  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
peak = 1385.6
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("bf16_tflops_sustained", peak)
dev = "cuda:0"
lib = _lib.load()
sd = synthetic.seeded_state_dict(0)
print("# TemporalUNet forward, precision %s; peak = %.1f TFLOP/s (sustained bf16, MEASURED_PEAKS.json)" % (prec, peak), file=out)
for rows in (256, 1024, 2048, 4096, 8192, 16384):
    m = TemporalUNet(os.path.join(tempfile.mkdtemp(), "m"), 7, 32, dev, dims=(32, 64, 128, 256, 512, 512), precision=prec)
    m.load_state_dict(sd)
    x = torch.randn(rows, 7, 50, device=dev)
    eps = m(x, 100)
    h = m.engine(rows)
    n = lib.edmp_unet_launches_per_forward(h)
    ms = np.zeros(n, dtype=np.float32)
    macs = np.zeros(n, dtype=np.float64)
    _lib.check(lib.edmp_unet_profile(h, ctypes.c_void_p(x.data_ptr()), 128, rows, 10, ms.ctypes.data_as(ctypes.c_void_p),
                                     macs.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(eps.data_ptr()), _lib.stream_ptr()),
               "edmp_unet_profile")
    kinds = {}
    for i in range(n):
        k = lib.edmp_unet_op_kernel(h, i).decode()
        kinds[k] = kinds.get(k, 0.0) + float(ms[i])
    tf = 2 * macs.sum() / (ms.sum() * 1e-3) / 1e12
    print("rows %6d: %8.3f ms/forward  %9.0f rows/s  %7.1f useful TFLOP/s  %5.1f %% of peak (x3 MMA work: %5.1f %%)  by kernel %s"
          % (rows, ms.sum(), rows / ms.sum() * 1e3, tf, 100 * tf / peak, 300 * tf / peak,
             {k: round(v, 3) for k, v in kinds.items()}), file=out)
    del m
    torch.cuda.empty_cache()

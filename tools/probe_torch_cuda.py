"""Where does the stock PyTorch-CUDA eager port spend its time? (UNet forward vs autograd guide, per step)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import guide_oracle as go, unet_oracle
dev = "cuda:0"
rows_per_guide = int(sys.argv[1]) if len(sys.argv) > 1 else 102
guides = bench.CONFIGS["c2"]["guides"]
sd = {k: v.to(dev) for k, v in bench.synthetic_state_dict().items()}
cfgs, scene, x_T, start, goal = bench.build_workload(guides, rows_per_guide)
rows = cfgs["total_batch_size"]
x = torch.randn(rows, 7, 50, device=dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for bm in (False, True):
    torch.backends.cudnn.benchmark = bm
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            ms = timeit(lambda: unet_oracle.unet_forward(sd, x, 100))
        print("rows %d cudnn.benchmark=%s tf32=%s: UNet forward %.2f ms" % (rows, bm, tf32, ms))
m = torch.tensor(cfgs["guidance_method"], dtype=torch.float32, device=dev).view(rows, 1, 1)
def grad():
    q = (0.3 * torch.randn(rows, 7, 48, device=dev)).requires_grad_(True)
    omin, omax = go.obstacle_aabbs(scene, cfgs["expansion"][:, 99], cfgs["clearance"][:, 99], rows=rows, device=dev)
    cost = torch.sum((1 - m) * go.iv_cost(q, omin, omax)) + torch.sum(m * go.sv_cost(q, start, goal, omin, omax))
    cost.backward()
    return q.grad
print("guide gradient (autograd): %.2f ms" % timeit(grad))
def aabbs():
    return go.obstacle_aabbs(scene, cfgs["expansion"][:, 99], cfgs["clearance"][:, 99], rows=rows, device=dev)
print("  of which obstacle_aabbs: %.2f ms" % timeit(aabbs))

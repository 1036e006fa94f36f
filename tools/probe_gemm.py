"""Launch one linear_fwd shape a few times (for ncu captures and quick timing)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vit3d_b200
from vit3d_b200._lib import PREC, call, ptr, stream

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=66560)
ap.add_argument("--n", type=int, default=2048)
ap.add_argument("--k", type=int, default=256)
ap.add_argument("--prec", default="bf16")
ap.add_argument("--act", type=int, default=1)
ap.add_argument("--res", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = "cuda:0"
lp = a.prec == "bf16"
adt = torch.bfloat16 if lp else torch.float32
x = (torch.randn(a.m, a.k, device=dev) * 0.5).to(adt)
w = torch.randn(a.n, a.k, device=dev) * 0.05
wl = w.to(torch.bfloat16) if lp else None
b = torch.randn(a.n, device=dev) * 0.01
res = torch.randn(a.m, a.n, device=dev) if a.res else None
yf = (not lp) or a.res
y = torch.empty(a.m, a.n, device=dev, dtype=torch.float32 if yf else torch.bfloat16)


def run():
    call("vit3d_linear_fwd", ptr(x), a.k, int(not lp), ptr(w), ptr(wl), ptr(b), ptr(res), ptr(y), int(yf), None, a.act,
         a.m, a.n, a.k, PREC[a.prec], stream())


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"M={a.m} N={a.n} K={a.k} {a.prec} act={a.act} res={a.res}: {ms*1e3:.1f} us  {2.0*a.m*a.n*a.k/ms/1e9:.1f} TFLOP/s")

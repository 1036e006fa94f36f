"""Time the attention kernels alone."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200
from vit3d_b200._lib import PREC, call, ptr, stream
ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=1024)
ap.add_argument("--heads", type=int, default=8)
ap.add_argument("--vis", type=int, default=1)
ap.add_argument("--bwd", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = "cuda:0"
S, A = 65, 256
qkv = (torch.randn(a.b, S, 3 * A, device=dev)).to(torch.bfloat16)
ctx = torch.empty(a.b, S, A, device=dev, dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
probs = torch.empty(a.b, a.heads, S, S, device=dev) if a.vis else None
def run():
    if a.bwd:
        call("vit3d_attn_bwd", ptr(ctx), ptr(qkv), ptr(dqkv), a.b, S, a.heads, A // a.heads, PREC["bf16"], stream())
    else:
        call("vit3d_attn_fwd", ptr(qkv), ptr(ctx), ptr(probs), a.b, S, a.heads, A // a.heads, PREC["bf16"], stream())
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
byts = a.b * S * (3 * A + A) * 2 + (a.b * a.heads * S * S * 4 if a.vis else 0)
if a.bwd: byts = a.b * S * (3 * A * 2 + A) * 2
print(f"B={a.b} heads={a.heads} vis={a.vis} bwd={a.bwd}: {ms*1e3:.1f} us, {byts/ms/1e6:.0f} GB/s algorithmic")

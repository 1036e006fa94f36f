"""Launch the attention forward kernel a few times (for ncu captures).  --heads 8|16|4  --vis 0|1|2
(1: packed probability rows of 65 floats, 2: rows padded to 72 floats - what the module path uses)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200  # noqa: F401
from vit3d_b200._lib import PREC, call, ptr, stream
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--heads", type=int, default=8)
ap.add_argument("--vis", type=int, default=2)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
B, S, H = a.batch, 65, 256
qkv = (torch.randn(B * S, 3 * H, device=dev) * 0.5).to(torch.bfloat16)
ctx = torch.empty(B * S, H, device=dev, dtype=torch.bfloat16)
probs = torch.empty(B, a.heads, S, 72 if a.vis == 2 else S, device=dev) if a.vis else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters + 1):
    if i == 1:
        e0.record()
    if a.vis == 2:
        call("vit3d_attn_fwd_padded", ptr(qkv), ptr(ctx), ptr(probs), 72, B, S, a.heads, H // a.heads, stream())
    else:
        call("vit3d_attn_fwd", ptr(qkv), ptr(ctx), ptr(probs), B, S, a.heads, H // a.heads, PREC["bf16"], stream())
e1.record()
torch.cuda.synchronize()
print(f"attention fwd B={B} heads={a.heads} vis={a.vis}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us")

"""Launch the fused MLP (+LN) kernel a few times (for ncu captures).  --pair 0|1"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200  # noqa: F401
from vit3d_b200._lib import call, lib, ptr, stream
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--d", type=int, default=2048)
ap.add_argument("--pair", type=int, default=1, help="0 one CTA per tile, 1 multicast pairs, 2 cta_group::2 pairs")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
M, H, d = a.batch * 65, 256, a.d
xn = (torch.randn(M, H, device=dev) * 0.8).to(torch.bfloat16)
w1 = (torch.randn(d, H, device=dev) / 16).to(torch.bfloat16)
w2 = (torch.randn(H, d, device=dev) / d ** 0.5).to(torch.float16)
b1 = torch.randn(d, device=dev) * 0.01
b2 = torch.randn(H, device=dev) * 0.01
x32 = torch.randn(M, H, device=dev)
y32 = torch.empty(M, H, device=dev)
yn = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
g = torch.ones(H, device=dev)
be = torch.zeros(H, device=dev)
lib().vit3d_set_tuning(5, a.pair)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters + 1):
    if i == 1:
        e0.record()
    call("vit3d_mlp_ln_fwd", ptr(xn), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(x32), ptr(y32), ptr(g), ptr(be), 1e-6,
         ptr(yn), M, H, d, stream())
e1.record()
torch.cuda.synchronize()
print(f"pair={a.pair} M={M} d={d}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us")

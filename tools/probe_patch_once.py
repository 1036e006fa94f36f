"""Launch the patch-embedding kernel a few times (for ncu captures)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200  # noqa: F401
from vit3d_b200._lib import PREC, call, lib, ptr, stream
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
B, H = a.batch, 256
vol = torch.randn(B, 1, 128, 128, 5, device=dev)
wp = torch.randn(H, 1, 16, 16, 5, device=dev) * 0.02
bp = torch.randn(H, device=dev) * 0.01
cls = torch.randn(1, 1, H, device=dev) * 0.02
pos = torch.randn(1, 65, H, device=dev) * 0.02
tok = torch.empty(B, 65, H, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters + 1):
    if i == 1:
        e0.record()
    call("vit3d_patch_embed_fwd", ptr(vol), ptr(wp), ptr(bp), ptr(cls), ptr(pos), ptr(tok), B, 128, 128, 5, 16, 16, 5, H,
         PREC["bf16"], None, 0, stream())
e1.record()
torch.cuda.synchronize()
print(f"patch embedding B={B}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us")

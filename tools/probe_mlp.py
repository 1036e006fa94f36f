"""Time the fused MLP kernel alone."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200
from vit3d_b200._lib import call, ptr, stream
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=66560)
ap.add_argument("--d", type=int, default=2048)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--timeline", type=int, default=0)
a = ap.parse_args()
dev = "cuda:0"
H = 256
xn = torch.randn(a.m, H, device=dev).to(torch.bfloat16)
w1 = (torch.randn(a.d, H, device=dev) / 16).to(torch.bfloat16)
w2 = (torch.randn(H, a.d, device=dev) / 45).to(torch.float16)
b1 = torch.randn(a.d, device=dev) * 0.1
b2 = torch.randn(H, device=dev) * 0.1
res = torch.randn(a.m, H, device=dev)
out = torch.empty(a.m, H, device=dev)
def run():
    call("vit3d_mlp_fwd", ptr(xn), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(res), ptr(out), a.m, H, a.d, stream())
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"fused MLP M={a.m} d={a.d}: {ms*1e3:.1f} us  {4.0*a.m*H*a.d/ms/1e9:.1f} TFLOP/s")

if a.timeline:
    import ctypes
    dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
    L = vit3d_b200._lib.lib()
    L.vit3d_debug_set_mlp_timeline.argtypes = [ctypes.c_void_p]
    L.vit3d_debug_set_mlp_timeline(dbg.data_ptr())
    run(); torch.cuda.synchronize()
    L.vit3d_debug_set_mlp_timeline(None)
    t = dbg.cpu().tolist()
    t0 = min(v for v in t if v > 0)
    rel = lambda v: (v - t0) if v > 0 else -1
    print("producer issue times (op: clk):", [rel(t[i]) for i in range(16)])
    print("MMA per op [start, after A/acc wait, after w_full, after issue]:")
    for op in range(14):
        print("  op", op, [rel(t[64 + op * 4 + k]) for k in range(4)])
    print("epilogue per chunk [enter, acc1_full seen, after tmem ld, after gelu, after a_empty, after arrive]:")
    for c in range(10):
        print("  c", c, [rel(t[320 + c * 6 + k]) for k in range(6)])

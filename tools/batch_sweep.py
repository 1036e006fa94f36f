"""BASELINE.json config 2: conf 5 inference on one B200, batch sweep 1..1024 (volumes resident in HBM, model(x) in
eval / no_grad, bf16 mode, vis=True, one CUDA-graph replay per step, CUDA events on the launching stream).

    python tools/batch_sweep.py [--conf 5] [--precision bf16] [--vis 1] [--json out.json]

Prints one line per batch size: ms per step (= latency of the batch), volumes/s, model TFLOP/s."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200
from oracle import vit3d_oracle as O
from vit3d_b200.graphs import GraphedInference
from vit3d_b200.models.modeling import VisionTransformer

ap = argparse.ArgumentParser()
ap.add_argument("--conf", type=int, default=5)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--vis", type=int, default=1)
ap.add_argument("--json", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = vit3d_b200.north_star_config(a.conf)
flops = O.fwd_flops_per_volume(cfg)
m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, vis=bool(a.vis), precision=a.precision)
m.load_state_dict(O.init_state_dict(cfg, seed=42))
m.to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > L2: written between timed replays of small batches
rows = []
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    x = O.synth_volumes(B, seed=42).to(dev)
    g = GraphedInference(m)
    x = g.input_like(x)            # the resident batch is the graph's input buffer
    for _ in range(3):
        g(x)
    torch.cuda.synchronize()
    iters = 50 if B <= 64 else 20
    small = B * 327680 < (126 << 20)          # inputs smaller than L2: flush it between iterations
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if small:
        for _ in range(iters):
            flush.fill_(1)
            e0.record()
            g(x)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    else:
        e0.record()
        for _ in range(iters):
            g(x)
        e1.record()
        torch.cuda.synchronize()
        tot = e0.elapsed_time(e1)
    ms = tot / iters
    vps = B / (ms * 1e-3)
    rows.append({"batch": B, "ms_per_step": ms, "volumes_per_s": vps, "model_tflops": vps * flops / 1e12,
                 "l2": "flushed between steps" if small else "inputs larger than L2"})
    print(f"B={B:5d}: {ms:8.3f} ms/step  {vps:10.0f} volumes/s  {vps * flops / 1e12:7.1f} TFLOP/s  ({rows[-1]['l2']})")
    del g
if a.json:
    json.dump({"conf": a.conf, "precision": a.precision, "vis": bool(a.vis), "rows": rows}, open(a.json, "w"), indent=1)

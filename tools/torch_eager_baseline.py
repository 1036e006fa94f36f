"""Second comparator of BASELINE.md section 3: the same architecture written with stock torch.nn modules and run by
PyTorch eager on the GPU (cuDNN Conv3d, cuBLAS Linear / matmul, native LayerNorm / softmax / GELU) - the "existing
GPU kernels" bar the sm_100a path has to beat.  fp32, TF32 and bf16 autocast; inference with the attention
probabilities kept (the reference's vis=True default) and a training step (forward + backward, dropout on).

    python tools/torch_eager_baseline.py [--conf 5] [--batches 64,256,1024] [--train-conf 18 --train-batch 256]

This file does not use the library or the oracle: it is a plain nn.Module ViT with the reference's shapes
(hidden 256, 65 tokens, patch 16x16x5, exact-erf GELU, pre-LN blocks, BCE-with-logits loss)."""
import argparse, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as TF

CONFS = {5: (2048, 6, 8), 9: (2048, 8, 16), 11: (3072, 4, 8), 18: (3072, 8, 16)}     # mlp width, layers, heads
FLOPS = {5: 1.0903e9, 9: 1.4397e9, 11: 1.0135e9, 18: 1.9850e9}                         # forward, per volume


class Block(nn.Module):
    def __init__(self, H, d, heads, p):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(H, eps=1e-6), nn.LayerNorm(H, eps=1e-6)
        self.q, self.k, self.v, self.o = (nn.Linear(H, H) for _ in range(4))
        self.fc1, self.fc2 = nn.Linear(H, d), nn.Linear(d, H)
        self.drop = nn.Dropout(p)
        self.heads = heads

    def forward(self, x):
        B, S, H = x.shape
        h = x
        y = self.n1(x)
        sp = lambda t: t.view(B, S, self.heads, H // self.heads).permute(0, 2, 1, 3)
        q, k, v = sp(self.q(y)), sp(self.k(y)), sp(self.v(y))
        probs = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(H // self.heads), dim=-1)
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(B, S, H)
        x = self.o(ctx) + h
        h = x
        y = self.fc2(self.drop(TF.gelu(self.fc1(self.n2(x)))))
        return self.drop(y) + h, probs


class ViT(nn.Module):
    def __init__(self, d, L, heads, H=256, p=0.1):
        super().__init__()
        self.patch = nn.Conv3d(1, H, kernel_size=(16, 16, 5), stride=(16, 16, 5))
        self.cls = nn.Parameter(torch.randn(1, 1, H) * 0.02)
        self.pos = nn.Parameter(torch.randn(1, 65, H) * 0.02)
        self.drop = nn.Dropout(p)
        self.blocks = nn.ModuleList(Block(H, d, heads, p) for _ in range(L))
        self.norm = nn.LayerNorm(H, eps=1e-6)
        self.head = nn.Linear(H, 1)

    def forward(self, x, labels=None, pos_weight=None):
        t = self.patch(x).flatten(2).transpose(-1, -2)
        t = self.drop(torch.cat((self.cls.expand(t.shape[0], -1, -1), t), dim=1) + self.pos)
        attn = []
        for b in self.blocks:
            t, pr = b(t)
            attn.append(pr)
        t = self.norm(t)
        logits = self.head(t[:, 0])
        if labels is not None:
            return TF.binary_cross_entropy_with_logits(logits.view(-1, 1).float(), labels.view(-1, 1), pos_weight=pos_weight)
        return logits, attn, t


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conf", type=int, default=5)
    ap.add_argument("--batches", default="64,256,1024")
    ap.add_argument("--train-conf", type=int, default=18)
    ap.add_argument("--train-batch", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    for mode in ("fp32", "tf32", "bf16 autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        ac = torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16"))
        m = ViT(*CONFS[a.conf]).to(dev).eval()
        for B in [int(b) for b in a.batches.split(",")]:
            x = torch.randn(B, 1, 128, 128, 5, device=dev) * 45.0
            with torch.no_grad(), ac:
                ms = timed(lambda: m(x), 10 if B >= 256 else 30)
            print(f"inference conf {a.conf} {mode:14s} B={B:5d}: {ms:8.3f} ms/step  {B / ms * 1e3:10.0f} volumes/s  "
                  f"{B / ms * 1e3 * FLOPS[a.conf] / 1e12:6.1f} TFLOP/s", flush=True)
        del m
        m = ViT(*CONFS[a.train_conf]).to(dev).train()
        opt = torch.optim.SGD(m.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-2)
        B = a.train_batch
        x = torch.randn(B, 1, 128, 128, 5, device=dev) * 45.0
        y = (torch.rand(B, device=dev) > 0.5).float()
        pw = torch.tensor(1.3, device=dev)

        def step():
            opt.zero_grad(set_to_none=True)
            with ac:
                loss = m(x, y, pw)
            loss.backward()
            opt.step()

        ms = timed(step, 5)
        print(f"training  conf {a.train_conf} {mode:14s} B={B:5d}: {ms:8.3f} ms/step  {B / ms * 1e3:10.0f} volumes/s  "
              f"{B / ms * 1e3 * 3 * FLOPS[a.train_conf] / 1e12:6.1f} TFLOP/s", flush=True)
        del m, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

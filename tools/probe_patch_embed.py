"""Times the patch embedding (TMA im2col TF32 GEMM + cls rows) alone at the conf-5 geometry, CUDA events over 20
launches on three rotating input batches (1 GB in total: larger than L2), 128-row tiles, 256-row tiles, clusters of 2 / 4 CTAs sharing the filter bank by TMA multicast.

    python tools/probe_patch_embed.py [--batch 1024]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vit3d_b200  # noqa: E402
from vit3d_b200._lib import lib  # noqa: E402
from vit3d_b200.models.modeling import Embeddings  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = vit3d_b200.get_config(16, 2048, 6, 256, 8)
    emb = Embeddings(cfg, 128)
    emb.precision = "bf16"
    emb.to(dev).eval()
    xs = [torch.randn(args.batch, 1, 128, 128, 5, device=dev) for _ in range(3)]
    nbytes = xs[0].numel() * 4 + args.batch * 65 * 256 * 4
    for tall, cluster in ((0, 0), (1, 0), (0, 2), (0, 4), (0, 0), (0, 4)):
        lib().vit3d_set_tuning(11, tall)
        lib().vit3d_set_tuning(12, cluster)
        with torch.no_grad():
            for i in range(5):
                emb(xs[i % 3])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                emb(xs[i % 3])
            e1.record()
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"patch embedding B={args.batch} tall={tall} cluster={cluster}: {us:7.1f} us  {nbytes / us / 1e3:6.0f} GB/s (volume read + token write)")
    lib().vit3d_set_tuning(11, 0)
    lib().vit3d_set_tuning(12, 0)


if __name__ == "__main__":
    main()

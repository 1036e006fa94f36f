"""Times every kernel of the fused BF16 training step alone (CUDA events, 20 launches after 5 warm-up, inputs far
larger than L2 in aggregate) at the conf-18 / batch-256 shapes of BASELINE.json config 4 and prints one line per
kernel with its algorithmic bytes / FLOPs against the measured peaks.

    python tools/probe_train_kernels.py [--batch 256] [--d 3072] [--heads 16] [--only name,name]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vit3d_b200  # noqa: E402
from vit3d_b200._lib import PREC, call, ptr, stream  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--d", type=int, default=3072)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--only", default="")
    ap.add_argument("--tune", default="", help="comma list of key=value tuning switches (include/vit3d.h VIT3D_TUNE_*)")
    args = ap.parse_args()
    for kv in [t for t in args.tune.split(",") if t]:
        k, v = kv.split("=")
        vit3d_b200._lib.lib().vit3d_set_tuning(int(k), int(v))
    dev = torch.device("cuda:0")
    B, S, H, d, heads = args.batch, 65, 256, args.d, args.heads
    M = B * S
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    bf, f32 = torch.bfloat16, torch.float32
    g = torch.Generator(device=dev).manual_seed(1)

    def rnd(*shape, dtype=bf, s=0.1):
        return (torch.randn(*shape, device=dev, generator=g) * s).to(dtype)

    xn, gy = rnd(M, H), rnd(M, H)
    x32, g32, dy32 = rnd(M, H, dtype=f32), rnd(M, H, dtype=f32), rnd(M, H, dtype=f32)
    out32 = torch.empty(M, H, device=dev)
    outb = torch.empty(M, H, device=dev, dtype=bf)
    wide, wide2, pre = rnd(M, d), torch.empty(M, d, device=dev, dtype=bf), rnd(M, d, s=1.0)
    w1, w2 = rnd(d, H), rnd(H, d)
    w1_t, w2_t = w1.t().contiguous(), w2.t().contiguous()
    wqkv, wqkv_t, wo = rnd(3 * H, H), rnd(H, 3 * H), rnd(H, H)
    qkv, dqkv = rnd(M, 3 * H, s=0.5), torch.empty(M, 3 * H, device=dev, dtype=bf)
    b1, b2, bq = rnd(d, dtype=f32), rnd(H, dtype=f32), rnd(3 * H, dtype=f32)
    gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dwA, dwB = torch.zeros(d, H, device=dev), torch.zeros(H, d, device=dev)
    dwq = [torch.zeros(H, H, device=dev) for _ in range(3)]
    dbv = [torch.zeros(H, device=dev) for _ in range(3)]
    db1 = torch.zeros(d, device=dev)
    dg, dbt, dbo = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    L = 8
    nel = [M * H] + [n for _ in range(L) for n in (M * d, M * H)]
    bits_all = torch.empty(sum(nel) // 8, device=dev, dtype=torch.uint8)
    bits_wide = bits_all[M * H // 8: M * H // 8 + M * d // 8]
    bits_h = bits_all[:M * H // 8]
    sc = 1.0 / 0.9
    st = stream()
    bfp = PREC["bf16"]

    K = {}

    def reg(name, fn, bytes_=0, flops=0):
        K[name] = (fn, bytes_, flops)

    reg("dropout_bits (all sites, L=8)", lambda: call("vit3d_dropout_bits", ptr(bits_all), len(nel), (C.c_uint * len(nel))(*range(len(nel))),
                                                     (C.c_longlong * len(nel))(*nel), 0.1, 1234, 1, None, st), bytes_=sum(nel) // 8)
    reg("ln256_fwd (+dropout)", lambda: call("vit3d_ln256_fwd", ptr(x32), ptr(bits_h), sc, ptr(out32), ptr(gam), ptr(bet), ptr(outb), None,
                                             ptr(mean), ptr(rstd), M, 1e-6, st), bytes_=M * H * (4 + 4 + 2))
    reg("qkv GEMM fwd", lambda: call("vit3d_linear_fwd", ptr(xn), H, 0, ptr(wqkv), ptr(wqkv), ptr(bq), None, ptr(dqkv), 0, None, 0, M,
                                     3 * H, H, bfp, st), bytes_=M * H * 2 * 4, flops=2 * M * 3 * H * H)
    reg("attention fwd (no probs)", lambda: call("vit3d_attn_fwd", ptr(qkv), ptr(outb), None, B, S, heads, H // heads, bfp, st),
        bytes_=M * H * 2 * 4)
    reg("out-proj + residual + LN (train)", lambda: call("vit3d_linear_res_train_fwd", ptr(xn), ptr(wo), ptr(b2), ptr(x32), ptr(out32),
                                                         None, 1.0, ptr(gam), ptr(bet), 1e-6, ptr(outb), ptr(mean), ptr(rstd), M, H, H, st),
        bytes_=M * H * (2 + 4 + 4 + 2), flops=2 * M * H * H)
    reg("fc1 + GELU + pre + dropout", lambda: call("vit3d_fc1_train_fwd", ptr(xn), ptr(w1), ptr(b1), ptr(wide2), ptr(wide), ptr(bits_wide), sc,
                                                   M, d, H, st), bytes_=M * H * 2 + 2 * M * d * 2, flops=2 * M * d * H)
    reg("fc2 + dropout + residual + LN", lambda: call("vit3d_linear_res_train_fwd", ptr(wide), ptr(w2), ptr(b2), ptr(x32), ptr(out32),
                                                      ptr(bits_h), sc, ptr(gam), ptr(bet), 1e-6, ptr(outb), ptr(mean), ptr(rstd), M, H, d, st),
        bytes_=M * d * 2 + M * H * (4 + 4 + 2), flops=2 * M * d * H)
    reg("wgrad fc2 [H,d]", lambda: call("vit3d_wgrad", ptr(gy), ptr(wide), ptr(dwB), None, None, 0, M, H, d, st),
        bytes_=M * d * 2 + M * H * 2, flops=2 * M * d * H)
    ws = torch.empty(vit3d_b200._lib.lib().vit3d_wgrad_ws_bytes(M, H, d) // 4, device=dev)
    bn_, sp_ = C.c_int(0), C.c_int(0)
    reg("wgrad fc2 partial tiles (no reduce)", lambda: call("vit3d_wgrad_partial", ptr(gy), ptr(wide), ptr(ws), M, H, d, C.byref(bn_),
                                                            C.byref(sp_), st), bytes_=M * d * 2 + M * H * 2, flops=2 * M * d * H)
    reg("mlp_bwd fused (dgrad fc2, x dact, dgrad fc1)", lambda: call("vit3d_mlp_bwd", ptr(gy), ptr(w2_t), ptr(w1_t), ptr(pre), ptr(wide2),
                                                                    ptr(out32), ptr(db1), M, H, d, st),
        bytes_=2 * M * d * 2 + M * H * 6, flops=4 * M * d * H)
    reg("dgrad fc2 (unfused)", lambda: call("vit3d_linear_fwd", ptr(gy), H, 0, ptr(w2_t), ptr(w2_t), None, None, ptr(wide2), 0, None, 0, M, d,
                                            H, bfp, st), bytes_=M * d * 2 + M * H * 2, flops=2 * M * d * H)
    reg("mul_colsum_bwd (unfused)", lambda: call("vit3d_mul_colsum_bwd", ptr(wide), ptr(pre), ptr(wide2), ptr(db1), M, d, st),
        bytes_=3 * M * d * 2)
    reg("dgrad fc1 (unfused)", lambda: call("vit3d_linear_fwd", ptr(wide), d, 0, ptr(w1_t), ptr(w1_t), None, None, ptr(out32), 1, None, 0, M,
                                            H, d, bfp, st), bytes_=M * d * 2 + M * H * 4, flops=2 * M * d * H)
    reg("wgrad fc1 [d,H]", lambda: call("vit3d_wgrad", ptr(wide), ptr(xn), ptr(dwA), None, None, 0, M, d, H, st),
        bytes_=M * d * 2 + M * H * 2, flops=2 * M * d * H)
    reg("ln256_bwd (+skip, bf16 copy, colsum)", lambda: call("vit3d_ln256_bwd", ptr(dy32), ptr(x32), ptr(mean), ptr(rstd), ptr(gam), ptr(g32),
                                                             None, 1.0, 0, ptr(out32), ptr(outb), ptr(dg), ptr(dbt), ptr(dbo), M, st),
        bytes_=M * H * (4 * 4 + 2))
    reg("wgrad out-proj [H,H]", lambda: call("vit3d_wgrad", ptr(gy), ptr(xn), ptr(dwq[0]), None, None, 0, M, H, H, st),
        bytes_=2 * M * H * 2, flops=2 * M * H * H)
    reg("dgrad out-proj", lambda: call("vit3d_linear_fwd", ptr(gy), H, 0, ptr(wo), ptr(wo), None, None, ptr(outb), 0, None, 0, M, H, H, bfp, st),
        bytes_=2 * M * H * 2, flops=2 * M * H * H)
    reg("attention bwd (+bias grads)", lambda: call("vit3d_attn_bwd_bias", ptr(xn), ptr(qkv), ptr(dqkv), ptr(dbv[0]), ptr(dbv[1]), ptr(dbv[2]),
                                                    B, S, heads, H // heads, st), bytes_=M * H * 2 * 7)
    reg("wgrad qkv [3H,H] (3 segments)", lambda: call("vit3d_wgrad", ptr(qkv), ptr(xn), ptr(dwq[0]), ptr(dwq[1]), ptr(dwq[2]), H, M, 3 * H, H, st),
        bytes_=M * H * 2 * 4, flops=2 * M * 3 * H * H)
    reg("dgrad qkv", lambda: call("vit3d_linear_fwd", ptr(qkv), 3 * H, 0, ptr(wqkv_t), ptr(wqkv_t), None, None, ptr(out32), 1, None, 0, M, H,
                                  3 * H, bfp, st), bytes_=M * H * (6 + 4), flops=2 * M * 3 * H * H)

    only = [s for s in args.only.split(",") if s]
    total = 0.0
    print(f"shapes: B={B} M={M} H={H} d={d} heads={heads}; peaks: HBM {peaks['hbm_gbs']:.0f} GB/s, bf16 {peaks['bf16_tflops']:.0f} TFLOP/s (burst)")
    for name, (fn, nb, fl) in K.items():
        if only and not any(o in name for o in only):
            continue
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        total += us
        line = f"{us:8.1f} us  {name:45s}"
        if nb:
            gbs = nb / us / 1e3
            line += f" {gbs:7.0f} GB/s ({gbs / peaks['hbm_gbs']:.2f} of HBM)"
        if fl:
            tf = fl / us / 1e6
            line += f" {tf:7.0f} TFLOP/s ({tf / peaks['bf16_tflops']:.2f} of bf16)"
        print(line)
    print(f"{total:8.1f} us  sum")


if __name__ == "__main__":
    main()

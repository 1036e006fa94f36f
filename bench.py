#!/usr/bin/env python
"""Headline benchmark: 3D-ViT (conf 5 by default) inference volumes/s on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --impl reference ...       # the reference's CPU path (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" = one forward pass of `model(x)` (eval, no_grad, reference
call pattern train_baseline_cv.py:79) over one batch of `--batch` synthetic volumes per GPU.
`value` = volumes/s with the batch resident in HBM; `e2e` = the same through the public module call
with pinned HOST input, H2D copy and D2H read of the logits inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ensemble volumes/sec (inference, fwd+bwd train) at 1/2/4/8 B200 vs CPU ref"
UNIT = "volumes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="conf5_infer",
                    choices=["conf5_infer", "conf18_train", "ensemble_infer"])
    ap.add_argument("--batch", type=int, default=0, help="volumes per GPU per step (0 = workload default)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--vis", type=int, default=1, help="materialise attention probabilities (reference default vis=True)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="volumes per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphs", type=int, default=1, help="replay the step from a CUDA graph (vit3d_b200.graphs)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [v.strip() for v in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(args, workload_cfgs, train, steps, warmup, sample):
    """Times the reference's algorithm on the host cores: the oracle's functional restatement of
    models/modeling.py (torch CPU fp32, all threads).  Bounded sample of the same workload."""
    import torch
    from oracle import vit3d_oracle as O
    torch.set_num_threads(os.cpu_count())
    cfgs = workload_cfgs
    sds = [O.init_state_dict(c, seed=42 + j) for j, c in enumerate(cfgs)]
    x = O.synth_volumes(sample, seed=42)
    y = O.synth_labels(sample)
    w = O.balanced_pos_weight(y)
    ens_sd = O.ensemble_state_dict(sds) if len(cfgs) > 1 else None

    def step():
        if train:
            O.vit_loss_and_grads(sds[0], cfgs[0], x, y, w)
        elif ens_sd is not None:
            with torch.no_grad():
                O.ensemble_forward(ens_sd, cfgs, x)
        else:
            with torch.no_grad():
                O.vit_forward(sds[0], cfgs[0], x)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps * 1e3


def workload(args):
    import vit3d_b200
    if args.workload == "conf5_infer":
        return dict(name="conf5 (d2048 L6 8x32 H256) inference, model(x) eval/no_grad", confs=[5], train=False,
                    batch=args.batch or 1024)
    if args.workload == "conf18_train":
        return dict(name="conf18 (d3072 L8 16x16 H256) training fwd+bwd", confs=[18], train=True,
                    batch=args.batch or 256)
    return dict(name="ensemble conf 5+9+11 + meta-classifier inference", confs=[5, 9, 11], train=False,
                batch=args.batch or 512)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args)
    import torch
    import vit3d_b200
    from oracle import vit3d_oracle as O
    cfgs = [vit3d_b200.north_star_config(c) for c in wl["confs"]]
    flops_per_vol = sum(O.fwd_flops_per_volume(c) for c in cfgs) * (3.0 if wl["train"] else 1.0)

    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(args.cpu_sample, wl["batch"])
        steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        v, ms = cpu_arm(args, cfgs, wl["train"], steps, warm, sample)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_step": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{steps} steps x {sample} volumes, oracle port of models/modeling.py, torch CPU fp32"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from vit3d_b200.models.modeling import TransformerEnsemble, VisionTransformer
    L = vit3d_b200._lib.lib()

    B = wl["batch"]
    members = []
    for j, c in enumerate(cfgs):
        m = VisionTransformer(c, 128, zero_head=True, num_classes=1, vis=bool(args.vis), precision=args.precision)
        m.load_state_dict(O.init_state_dict(c, seed=42 + j))
        members.append(m)
    model = members[0] if len(members) == 1 else TransformerEnsemble(*members, in_features=1)
    model.to(dev)
    train = wl["train"]
    model.train(train)
    x_host = O.synth_volumes(B, seed=42 + (rank if len(members) == 1 else 0)).pin_memory()
    y_host = O.synth_labels(B).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = y_host.to(dev)
    opt = reducer = sharded = None
    if train:
        # the reference's optimizer (train_baseline_cv.py:111-114) as one fused launch over a flat arena,
        # gradients all-reduced per encoder Block while backward is still running
        from vit3d_b200.dist import GradReducer, global_pos_weight
        from vit3d_b200.optim import FusedSGD
        opt = FusedSGD(model.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-2)
        reducer = GradReducer(model, arena=opt.arena) if (world > 1 and not args.graphs) else None
    elif len(members) > 1:
        from vit3d_b200.dist import ShardedEnsemble
        sharded = ShardedEnsemble(model, costs=[O.fwd_flops_per_volume(c) for c in cfgs])

    graphed = None
    if args.graphs and train:
        from vit3d_b200.graphs import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, warmup=2, data_parallel=world > 1)
        reducer = None
    elif args.graphs and not train and sharded is None:
        from vit3d_b200.graphs import GraphedInference
        graphed = GraphedInference(model)
        # the resident batch lives in the graph's own input buffer (no device-to-device copy per replay); the
        # end-to-end loop below passes its double-buffered H2D targets, which the call copies in
        x_dev = graphed.input_like(x_dev)

    def step_dev(x, y):
        if graphed is not None:
            if train:
                return graphed(x, y, global_pos_weight(y) if world > 1 else O.balanced_pos_weight(y_host))
            out = graphed(x)
            return out[0] if isinstance(out, tuple) else out
        if train:
            pw = global_pos_weight(y) if world > 1 else O.balanced_pos_weight(y_host)
            if reducer is not None:
                reducer.prepare()
            else:
                opt.zero_grad()
            loss = model(x, y, pw)
            loss.backward()
            if reducer is not None:
                reducer.finish(scale=False)
            opt.step(grad_scale=1.0 / world)
            return loss
        if sharded is not None:
            return sharded(x)
        with torch.no_grad():
            out = model(x)
        return out[0] if isinstance(out, tuple) else out

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.vit3d_launch_count()
        r0 = graphed.replays if graphed is not None else 0
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        n1 = L.vit3d_launch_count()
        if graphed is not None:     # kernels replayed from the captured graph (counted once, at capture)
            n1 += (graphed.replays - r0) * graphed.launches_per_replay
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            dist.barrier()
        return ms, n1 - n0

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # timing rule: at least 3 untimed warm-up steps, whatever was asked for (the JSON line reports what was run)
    args.warmup = max(3, args.warmup)
    ms_dev, launches = timed(lambda: step_dev(x_dev, y_dev), args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None

    res_host = torch.empty((B, 1) if not train else (), dtype=torch.float32).pin_memory()

    # End-to-end: every step copies ITS batch from pinned host memory (H2D on a copy stream, double
    # buffered so the copy of step s+1 overlaps the compute of step s), runs the public module call and
    # reads the result back to the host (D2H) before the step counts as done.
    copy_stream = torch.cuda.Stream()
    xbuf = [torch.empty_like(x_dev) for _ in range(2)]
    ybuf = [torch.empty_like(y_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(s):
        b = s & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])          # the compute that read this buffer two steps ago is done
            xbuf[b].copy_(x_host, non_blocking=True)
            if train:
                ybuf[b].copy_(y_host, non_blocking=True)
            ready[b].record(copy_stream)

    def run_e2e(steps):
        main = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(main)
        issue_copy(0)
        for s in range(steps):
            b = s & 1
            main.wait_event(ready[b])
            out = step_dev(xbuf[b], ybuf[b] if train else None)
            consumed[b].record(main)
            if s + 1 < steps:
                issue_copy(s + 1)
            res_host.copy_(out.detach().reshape(res_host.shape), non_blocking=True)
            main.synchronize()

    def timed_e2e(steps, warm):
        run_e2e(warm)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
            dist.barrier()
        return ms

    ms_e2e = timed_e2e(args.steps, max(3, args.warmup // 2))

    # N2 variant of the end-to-end number (extra, not the contract's `e2e`): the same volumes cross PCIe as
    # uint8 (they ARE 8-bit images minus a mean) and (u8 - mean) runs on the device.
    e2e_u8 = None
    if not train and len(members) == 1:
        u8_host, mean = O.synth_volumes_u8(B, seed=42 + rank)
        u8_host = u8_host.pin_memory()
        model.input_mean = mean
        x_keep, xbuf_keep = x_host, xbuf
        x_host = u8_host
        xbuf = [torch.empty(u8_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        ms_u8 = timed_e2e(args.steps, max(3, args.warmup // 2))
        x_host, xbuf = x_keep, xbuf_keep
        e2e_u8 = {"value": world * B * args.steps / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(B * 81920),
                  "d2h_bytes_per_step": int(res_host.numel() * 4), "ms_per_step": ms_u8 / args.steps,
                  "note": "volumes shipped as uint8 + mean (N2 input path), (u8 - mean) on the device"}

    units = B if len(members) > 1 else world * B          # the sharded ensemble splits ONE batch over the ranks
    value = units * args.steps / (ms_dev * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (fc1/fc2 GEMM pair = 75-82 % of the FLOPs), timed alone
    roof = None
    cpu_base = None
    if rank == 0:
        roof = roofline_probe(args, cfgs[0], B, dev, train)
        if not args.no_cpu_baseline:
            sample = min(args.cpu_sample, B)
            v, ms = cpu_arm(args, cfgs, train, 3, 1, sample)
            cpu_base = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": f"3 steps x {sample} volumes of the same workload, oracle port of models/modeling.py, "
                                  f"torch CPU fp32, {os.cpu_count()} threads"}
    if rank == 0:
        peaks = load_peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if len(members) > 1 else "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "global_batch": units, "vis": bool(args.vis),
                       "precision": args.precision, "cuda_graph": graphed is not None, "l2": "inputs larger than L2 (batch of fp32 volumes = %.0f MB)" % (B * 327680 / 1e6),
                       "parallelism": (f"dp{world}: batch sharded, fwd+bwd replayed from a CUDA graph, one NCCL all-reduce of the flat gradient arena, fused SGD step"
                                       if train else (f"{world} ranks: (member, batch-slice) work list balanced by FLOPs, all-gather of member logits, meta-head on every rank"
                                                      if len(members) > 1 else f"dp{world} (independent volumes, no data-path collective)"))},
            "model_tflops": value * flops_per_vol / 1e12,
            "model_frac_of_bf16_sustained": value * flops_per_vol / 1e12 / (world * peaks["bf16_tflops_sustained"]),
            "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * 327680 + (B * 4 if train else 0)),
                    "d2h_bytes_per_step": int(res_host.numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "e2e_u8": e2e_u8,
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _time_launches(fn, iters=10, warm=3):
    import torch
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures under profiles/
# (only valid for the captured shape: conf 5, batch 1024, bf16)
NCU_TRAFFIC = {"mlp_ln": 162.7e6, "fc1_gelu": 257.2e6}


def roofline_probe(args, cfg, B, dev, train=False):
    """Times the dominant kernel alone with CUDA events on the launching stream.

    Inference in BF16 mode: the fused MLP block (fc1 -> GELU -> fc2 -> + residual -> next LayerNorm,
    vit3d_mlp_ln_fwd), 44 % of the step and 75-82 % of the model's FLOPs: tensor bound, 4*S*H*d FLOP per volume
    per launch.  Otherwise (training / other precisions): the fc1 + GELU GEMM.  `others` carries the HBM-bound
    kernels of a layer against the measured copy bandwidth."""
    import torch
    from vit3d_b200 import _lib
    from vit3d_b200._lib import PREC, call, ptr, stream
    peaks = load_peaks()
    L = _lib.lib()
    M, H, d = B * 65, cfg.hidden_size, cfg.transformer["mlp_dim"]
    heads = cfg.transformer["num_heads"]
    prec = args.precision
    lp = prec == "bf16"
    adt = torch.bfloat16 if lp else torch.float32
    xn = (torch.randn(M, H, device=dev) * 0.5).to(adt)
    w1 = torch.randn(d, H, device=dev) * 0.05
    b1 = torch.randn(d, device=dev) * 0.01
    captured = (M == 66560 and d == 2048 and H == 256 and lp)
    others = []
    fused = (not train) and lp and H == 256 and bool(L.vit3d_mlp_ln_supported(M, H, d))
    if fused:
        w1l = w1.to(torch.bfloat16)
        w2h = (torch.randn(H, d, device=dev) / d ** 0.5).to(torch.float16)
        b2 = torch.randn(H, device=dev) * 0.01
        x32 = torch.randn(M, H, device=dev)
        y32 = torch.empty(M, H, device=dev)
        yn = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
        g, be = torch.ones(H, device=dev), torch.zeros(H, device=dev)
        ms = _time_launches(lambda: call("vit3d_mlp_ln_fwd", ptr(xn), ptr(w1l), ptr(b1), ptr(w2h), ptr(b2), ptr(x32), ptr(y32),
                                         ptr(g), ptr(be), 1e-6, ptr(yn), M, H, d, stream()))
        flops = 4.0 * M * H * d
        ach = flops / (ms * 1e-3) / 1e12
        roof = {"kernel": "fused MLP block: fc1 + GELU + fc2 + residual + LayerNorm (vit3d_mlp_ln_fwd, tc_mlp2_kernel, "
                          "M=%d H=%d d=%d, tcgen05)" % (M, H, d),
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"],
                "traffic": NCU_TRAFFIC["mlp_ln"] if captured else None, "traffic_unit": "bytes per launch",
                "algorithmic_bytes": float(M * H * (2 + 4 + 4 + 2) + 2 * H * d * 2),
                "algorithmic_flops": flops, "ms_per_launch": ms, "peak_source": peaks.get("source"),
                "how": "kernel timed alone, 10 launches, CUDA events (burst peak applies)"}
        # HBM-bound kernels of the same layer (bytes = what the kernel must read + write once)
        hbm = peaks["hbm_gbs"]
        qkv = (torch.randn(M, 3 * H, device=dev) * 0.5).to(torch.bfloat16)
        ctx = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
        probs = torch.empty(B, heads, 65, 65, device=dev) if args.vis else None
        t = _time_launches(lambda: call("vit3d_attn_fwd", ptr(qkv), ptr(ctx), ptr(probs), B, 65, heads, H // heads,
                                        PREC["bf16"], stream()))
        nb = M * 4 * H * 2 + (B * heads * 65 * 65 * 4 if args.vis else 0)
        others.append({"kernel": "attention forward (attn_fwd_tc_kernel, vis=%s)" % bool(args.vis), "bound": "hbm",
                       "achieved": nb / (t * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nb / (t * 1e-3) / 1e9 / hbm,
                       "ms_per_launch": t})
        wo = (torch.randn(H, H, device=dev) / 16).to(torch.bfloat16)
        t = _time_launches(lambda: call("vit3d_linear_ln_fwd", ptr(ctx), ptr(wo), ptr(b2), ptr(x32), ptr(y32), ptr(g), ptr(be),
                                        1e-6, ptr(yn), None, None, M, H, H, stream()))
        nb = M * H * (2 + 4 + 4 + 2)
        others.append({"kernel": "out-projection + residual + LayerNorm (tc_gemm_res_kernel)", "bound": "hbm",
                       "achieved": nb / (t * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nb / (t * 1e-3) / 1e9 / hbm,
                       "ms_per_launch": t})
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        t = _time_launches(lambda: call("vit3d_ln_fwd", ptr(x32), ptr(g), ptr(be), ptr(yn), 1, ptr(mean), ptr(rstd), M, H, 1e-6,
                                        stream()))
        nb = M * H * 6
        others.append({"kernel": "LayerNorm fp32 -> bf16 (ln_fwd_kernel)", "bound": "hbm", "achieved": nb / (t * 1e-3) / 1e9,
                       "peak": hbm, "unit": "GB/s", "frac": nb / (t * 1e-3) / 1e9 / hbm, "ms_per_launch": t})
    # the fc1 + GELU GEMM (training forward; the inference path when the fused block is unavailable)
    wl = w1.to(torch.bfloat16) if lp else None
    y = torch.empty(M, d, device=dev, dtype=adt)
    tc = bool(L.vit3d_tc_supported(PREC[prec], M, d, H))
    ms1 = _time_launches(lambda: call("vit3d_linear_fwd", ptr(xn), H, int(adt == torch.float32), ptr(w1), ptr(wl), ptr(b1), None,
                                      ptr(y), int(adt == torch.float32), None, 1, M, d, H, PREC[prec], stream()))
    flops1 = 2.0 * M * d * H
    fc1 = {"kernel": "fc1+GELU GEMM (vit3d_linear_fwd, M=%d N=%d K=%d, %s)" % (M, d, H, "tcgen05" if tc else "fp32 FMA"),
           "bound": "tensor", "achieved": flops1 / (ms1 * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
           "frac": flops1 / (ms1 * 1e-3) / 1e12 / peaks["bf16_tflops"],
           "traffic": NCU_TRAFFIC["fc1_gelu"] if captured else None, "traffic_unit": "bytes per launch",
           "algorithmic_bytes": float(M * H * 2 + d * H * 2 + M * d * 2) if lp else None, "ms_per_launch": ms1,
           "peak_source": peaks.get("source"), "how": "kernel timed alone, 10 launches, CUDA events (burst peak applies)"}
    if not fused:
        return fc1
    others.append(fc1)
    roof["others"] = others
    return roof


if __name__ == "__main__":
    main()

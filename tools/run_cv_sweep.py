"""BASELINE.json config 5: the train_baseline_cv.py sweep (18 configurations x 5 CV folds = 90 independent training jobs of 100
steps at batch 4, validation every 24 steps on 18 volumes) packed on the GPUs of one box - one process per GPU, jobs
dealt out by dist.pack_jobs (longest first), several jobs in flight per GPU (workflow.run_packed_sweep).  Replicas
only: no collective on the data path.  Synthetic volumes of the reference's CV split sizes (72 train / 18 validation).

    python tools/run_cv_sweep.py [--configs 18] [--folds 5] [--steps 100] [--concurrent 6]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_cv_sweep.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vit3d_b200  # noqa: E402
from vit3d_b200 import workflow as W  # noqa: E402
from vit3d_b200.dist import pack_jobs  # noqa: E402


def synth_fold(fold, n_train=72, n_val=18):
    g = torch.Generator().manual_seed(1000 + fold)
    def vols(n):
        u8 = (66 + 45 * torch.randn(n, 1, 128, 128, 5, generator=g)).round().clamp(0, 255)
        return (u8 - u8.mean()).float().pin_memory()
    ty = torch.randint(0, 2, (n_train,), generator=g).float()
    vy = torch.randint(0, 2, (n_val,), generator=g).float()
    vy[0], vy[1] = 0.0, 1.0
    return vols(n_train), ty.pin_memory(), vols(n_val), vy


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=18)
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--concurrent", type=int, default=6)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    jobs = [(c, f) for c in range(1, args.configs + 1) for f in range(args.folds)]
    costs = []
    for c, _ in jobs:
        cfg = vit3d_b200.north_star_config(c)
        costs.append(cfg.transformer["num_layers"] * (4 * 256 + 2 * cfg.transformer["mlp_dim"]))     # ~ FLOPs per token
    mine = [jobs[i] for i in pack_jobs(costs, world)[rank]]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = W.run_packed_sweep(mine, synth_fold, steps=args.steps, batch=args.batch, concurrent=args.concurrent,
                             device=f"cuda:{local}")
    wall = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([wall], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t)
    if rank == 0:
        print(json.dumps({"workload": f"CV sweep: {args.configs} configurations x {args.folds} folds = {len(jobs)} jobs, {args.steps} steps at batch "
                                      f"{args.batch}, validation every 24 steps on 18 volumes", "n_gpus": world, "jobs": len(jobs),
                          "wall_s": wall, "jobs_per_s": len(jobs) / wall, "train_steps_per_s": len(jobs) * args.steps / wall,
                          "volumes_per_s": len(jobs) * args.steps * args.batch / wall, "concurrent_jobs_per_gpu": args.concurrent,
                          "rank0": {k: v for k, v in res.items() if k != "results"}, "collective": None,
                          "rank0_results_sample": res["results"][:3]}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

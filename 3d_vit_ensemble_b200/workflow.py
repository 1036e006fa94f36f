"""Host-side helpers for the "next" rows of SURVEY.md §8(f) - the callers on either side of the hot path.

N3  validation   `validate()` replaces the per-sample loop of train_baseline_cv.py:64-101, which runs the
                 forward TWICE per sample at batch 1 (:79-80): batches of volumes, ONE forward each, logits
                 and the cls feature taken from the same call; metrics stay on the host (sklearn in the
                 reference; plain tensors here).
N4  checkpoints  `ensemble_from_checkpoints()` builds a TransformerEnsemble from baseline state_dict files
                 the way train_ensemble_*.py intends to (train_ensemble_whole_dataset.py:50-52) without its
                 defects (it passes the RETURN VALUE of load_state_dict to the ensemble, uses
                 in_features=3 with 1-logit members, and looks for the files one directory too high);
                 `save_training_state()/load_training_state()` add the resume state (optimizer, scheduler,
                 step, dropout step) that the reference never saves (train_baseline_cv.py:128-134).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import functional as F
from .config import get_config, north_star_config, parameters_config
from .vit import TransformerEnsemble, VisionTransformer


# ----------------------------------------------------------------------------- N3
@torch.no_grad()
def validate(model: torch.nn.Module, volumes: torch.Tensor, labels: Optional[torch.Tensor] = None,
             batch_size: int = 256, graphs: bool = True) -> Dict[str, torch.Tensor]:
    """Batched single-forward validation.

    volumes: (N,1,128,128,5) fp32 or uint8 (N2 path, uses model.input_mean) on the host or the device.
    Returns host tensors: logits (N,), probs = sigmoid(logits) (N,), predicted = probs > 0.5 (N,),
    features (N, H) = the cls row of the encoder output (train_baseline_cv.py:80), and - with labels -
    accuracy, sensitivity, specificity, balanced accuracy (train_baseline_cv.py:94-96 without sklearn).
    For a TransformerEnsemble the model output already is the probability and there are no features."""
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    is_ens = isinstance(model, TransformerEnsemble)
    runner = model
    if graphs and dev.type == "cuda" and not is_ens:
        from .graphs import GraphedInference
        runner = GraphedInference(model)
    logits, feats = [], []
    N = volumes.shape[0]
    for i in range(0, N, batch_size):
        xb = volumes[i:i + batch_size].to(dev, non_blocking=True)
        if graphs and not is_ens and xb.shape[0] != batch_size and i > 0:
            runner_out = model(xb)                  # ragged tail: one eager call instead of a second capture
        else:
            runner_out = runner(xb)
        if is_ens:
            logits.append(runner_out.reshape(-1).float().cpu())
        else:
            lg, _, enc = runner_out
            logits.append(lg.reshape(-1).float().cpu())
            feats.append(enc[:, 0].float().cpu())
    z = torch.cat(logits) if logits else torch.zeros(0)
    out: Dict[str, torch.Tensor] = {}
    if is_ens:
        out["probs"] = z
    else:
        out["logits"] = z
        out["probs"] = torch.sigmoid(z)
        out["features"] = torch.cat(feats) if feats else torch.zeros(0, 0)
    out["predicted"] = (out["probs"] > 0.5).long()
    if labels is not None:
        y = labels.reshape(-1).long().cpu()
        p = out["predicted"]
        tp = int(((p == 1) & (y == 1)).sum()); tn = int(((p == 0) & (y == 0)).sum())
        fp = int(((p == 1) & (y == 0)).sum()); fn = int(((p == 0) & (y == 1)).sum())
        sens = tp / max(1, tp + fn)
        spec = tn / max(1, tn + fp)
        out["accuracy"] = torch.tensor((tp + tn) / max(1, y.numel()))
        out["sensitivity"] = torch.tensor(sens)
        out["specificity"] = torch.tensor(spec)
        out["balanced_accuracy"] = torch.tensor(0.5 * (sens + spec))
    model.train(was_training)
    return out


# ----------------------------------------------------------------------------- N4
def build_baseline(conf: int, *, as_shipped: bool = False, img_size: int = 128, vis: bool = True,
                   precision: Optional[str] = None) -> VisionTransformer:
    """One of the 18 baseline ViTs.  as_shipped=False: the README table (hidden 256 = head-dim x heads);
    as_shipped=True: what the reference's tools.parameters_config really returns (tools.py:60-80)."""
    cfg = get_config(*parameters_config(conf)) if as_shipped else north_star_config(conf)
    return VisionTransformer(cfg, img_size, zero_head=True, num_classes=1, vis=vis, precision=precision)


def ensemble_from_checkpoints(paths: Sequence[str], confs: Optional[Sequence[int]] = None, *, configs=None,
                              as_shipped: bool = False, device="cuda", precision: Optional[str] = None,
                              img_size: int = 128) -> TransformerEnsemble:
    """TransformerEnsemble of baseline members restored from their `torch.save(model.state_dict())` files
    (reference or this package: same keys and shapes).  Members output one logit each, so in_features=1.
    `confs`: configuration ids (README table, or as shipped); `configs`: explicit config objects instead."""
    specs = list(configs) if configs is not None else list(confs or [])
    if len(paths) != len(specs):
        raise ValueError("one checkpoint path per configuration")
    members = []
    for path, spec in zip(paths, specs):
        if configs is not None:
            m = VisionTransformer(spec, img_size, zero_head=True, num_classes=1, precision=precision)
        else:
            m = build_baseline(spec, as_shipped=as_shipped, precision=precision)
        sd = torch.load(path, map_location="cpu")
        missing, unexpected = m.load_state_dict(sd, strict=True)
        members.append(m)                          # the MODULE (the reference appends load_state_dict's return value)
    return TransformerEnsemble(*members, in_features=1).to(device)


def save_training_state(path: str, model: torch.nn.Module, optimizer=None, scheduler=None, step: int = 0,
                        extra: Optional[dict] = None) -> None:
    """Everything needed to resume: weights, optimizer buffers (flat arena for the fused optimizers),
    scheduler, global step and the dropout step counter."""
    state = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "step": int(step),
             "dropout_step": int(F._STATE["step"]), "extra": extra or {}}
    if optimizer is not None:
        if hasattr(optimizer, "arena"):
            state["optimizer"] = {"kind": type(optimizer).__name__, "steps": optimizer._steps,
                                  "param_groups": [{k: v for k, v in g.items() if k != "params"}
                                                   for g in optimizer.param_groups]}
            for name in ("momentum_buffer", "exp_avg", "exp_avg_sq"):
                buf = getattr(optimizer, name, None)
                if buf is not None:
                    state["optimizer"][name] = buf.detach().cpu()
        else:
            state["optimizer"] = optimizer.state_dict()
    if scheduler is not None:
        state["scheduler"] = scheduler.state_dict()
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)


def load_training_state(path: str, model: torch.nn.Module, optimizer=None, scheduler=None) -> int:
    """Inverse of save_training_state; returns the global step."""
    state = torch.load(path, map_location="cpu")
    model.load_state_dict(state["model"])
    F._STATE["step"] = int(state.get("dropout_step", 0))
    F.invalidate_weight_shadows()
    if optimizer is not None and "optimizer" in state:
        o = state["optimizer"]
        if hasattr(optimizer, "arena"):
            optimizer._steps = int(o.get("steps", 0))
            for g, sg in zip(optimizer.param_groups, o.get("param_groups", [])):
                g.update(sg)
            for name in ("momentum_buffer", "exp_avg", "exp_avg_sq"):
                if name in o and getattr(optimizer, name, None) is not None:
                    getattr(optimizer, name).copy_(o[name])
        else:
            optimizer.load_state_dict(o)
    if scheduler is not None and "scheduler" in state:
        scheduler.load_state_dict(state["scheduler"])
    return int(state.get("step", 0))


# ----------------------------------------------------------------------------- CV / bootstrap sweep (replicas only)
class _SweepJob:
    """One (configuration, fold) training run of train_baseline_cv.py:269-278: SGD(lr 1e-4, momentum .9, wd 1e-2) under the
    warm-up-cosine schedule (:111-119), `steps` optimizer steps at batch `batch`, validation every `eval_every` steps."""

    def __init__(self, conf, fold, device, precision, batch, seed):
        from .graphs import GraphedTrainStep
        from .optim import FusedSGD
        self.conf, self.fold = conf, fold
        torch.manual_seed(seed)                                   # tools.set_seed(args) per fold (train_baseline_cv.py:272)
        self.model = build_baseline(conf, precision=precision, vis=False).to(device)
        self.model.train()
        self.opt = FusedSGD(self.model.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-2)
        self.step_fn = GraphedTrainStep(self.model, self.opt, warmup=1)
        self.stream = torch.cuda.Stream(device=device)
        self.steps_done = 0
        self.best = -1.0
        self.losses = []


def warmup_cosine_lr(step: int, base_lr: float = 1e-4, warmup_steps: int = 1000, t_total: int = 100) -> float:
    """utils/scheduler.py:46-63 (WarmupCosineSchedule) as a host scalar: with the scripts' defaults (warm-up 1000 steps,
    100 steps in total) the learning rate never leaves the linear warm-up."""
    import math
    if step < warmup_steps:
        return base_lr * float(step) / float(max(1.0, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, t_total - warmup_steps))
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * 2.0 * 0.5 * progress)))


def run_packed_sweep(jobs: Sequence[Tuple[int, int]], data, *, steps: int = 100, batch: int = 4, eval_every: int = 24,
                     concurrent: int = 6, device="cuda", precision: str = "bf16", seed: int = 42) -> Dict[str, float]:
    """The CV sweep of train_baseline_cv.py (18 configurations x 5 folds, BASELINE.json config 5) as independent jobs
    packed on ONE GPU (one process per GPU takes its share of the job list from dist.pack_jobs: replicas only, no
    collective).  A batch-4 step is 260 token rows - three 128-row tiles on a 148-SM GPU - so `concurrent` jobs are
    kept in flight, each on its own stream, each step one CUDA-graph replay of the fused training step; the job
    streams interleave on the GPU.

    jobs: (configuration id, fold) pairs.  data(fold) -> (train_x (N,1,128,128,5), train_y (N,), val_x, val_y) host
    tensors (pinned for speed).  Returns wall-clock numbers: jobs/s, optimizer steps/s, time spent building models
    and capturing graphs, and per-job results (last loss, best validation balanced accuracy)."""
    import time
    from .dist import batch_pos_weight
    dev = torch.device(device)
    t0 = time.perf_counter()
    pending = list(jobs)
    active: List[_SweepJob] = []
    done = []
    setup_s = 0.0
    total_steps = 0
    cache = {}

    def fold_data(fold):
        if fold not in cache:
            cache[fold] = data(fold)
        return cache[fold]

    while pending or active:
        while pending and len(active) < concurrent:
            conf, fold = pending.pop(0)
            ts = time.perf_counter()
            active.append(_SweepJob(conf, fold, dev, precision, batch, seed + fold))
            setup_s += time.perf_counter() - ts
        for job in list(active):
            tx, ty, vx, vy = fold_data(job.fold)
            n = tx.shape[0]
            i0 = (job.steps_done * batch) % max(1, n - batch + 1)
            xb, yb = tx[i0:i0 + batch], ty[i0:i0 + batch].float()
            with torch.cuda.stream(job.stream):
                job.step_fn.set_lr(warmup_cosine_lr(job.steps_done + 1))
                loss = job.step_fn(xb.to(dev, non_blocking=True), yb.to(dev, non_blocking=True), batch_pos_weight(yb))
            job.steps_done += 1
            total_steps += 1
            if job.steps_done % eval_every == 0 or job.steps_done == steps:
                with torch.cuda.stream(job.stream):
                    res = validate(job.model, vx, vy, batch_size=max(1, vx.shape[0]), graphs=False)
                    job.model.train()
                    job.best = max(job.best, float(res["balanced_accuracy"]))
                    job.losses.append(float(loss))
            if job.steps_done >= steps:
                job.stream.synchronize()
                done.append({"conf": job.conf, "fold": job.fold, "loss": job.losses[-1] if job.losses else float("nan"),
                             "best_balanced_accuracy": job.best})
                active.remove(job)
                del job
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    F._STATE["step_dev"] = None
    return {"jobs": len(done), "wall_s": wall, "jobs_per_s": len(done) / wall, "steps_per_s": total_steps / wall,
            "setup_s": setup_s, "train_steps_per_s_excl_setup": total_steps / max(1e-9, wall - setup_s),
            "concurrent": concurrent, "results": done}

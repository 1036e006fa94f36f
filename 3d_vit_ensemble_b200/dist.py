"""Multi-GPU execution of the path on one 8xB200 box: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch; gloo in the CPU tests).  SURVEY.md §8e.

  * inference of one ViT      - volumes are independent: shard the batch, no data-path collective.
  * data-parallel training    - `GradReducer`: parameters' gradients live in ONE flat fp32 arena; a bucket
                                (one encoder Block, the embeddings, the head) is all-reduced as soon as
                                autograd has produced its last gradient, on a side stream, overlapping the
                                rest of backward.  `global_pos_weight` makes the per-batch class weight
                                (train_baseline_cv.py:168-169) that of the GLOBAL batch.
  * stacking ensemble         - `ShardedEnsemble`: the (member, batch-slice) work list is cut batch-major - rank r
                                runs every member on its 1/world_size of the batch (equal cost whatever the
                                members' FLOPs, 3 members need not divide 2/4/8 GPUs, and a rank ships only its
                                own slice of a host batch) - the members of a rank run concurrently on side
                                streams, one all-gather of the member logits rebuilds the (B, m) matrix
                                everywhere and the 3->1 meta-classifier runs redundantly.
  * CV / bootstrap sweep      - `pack_jobs`: independent jobs, replicas only, no collective.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ----------------------------------------------------------------------------- DP training
def class_pos_weight(n_pos: float, n: float) -> float:
    """The scalar the reference passes as `pos_weight` (train_baseline_cv.py:168-169):
    `sklearn.utils.class_weight.compute_class_weight('balanced', classes=unique(y), y=y)` gives
    n / (n_classes_present * count_c) per class; the script takes entry [1] when both classes are in
    the batch - n / (2 * n_pos) - and entry [0] otherwise, which is n / (1 * n) = 1.0."""
    if n_pos <= 0 or n_pos >= n:
        return 1.0
    return float(n) / (2.0 * float(n_pos))


def batch_pos_weight(labels: torch.Tensor) -> torch.Tensor:
    """`class_pos_weight` of one batch of 0/1 labels as the 0-dim float64 tensor the scripts build."""
    y = labels.reshape(-1)
    return torch.tensor(class_pos_weight(float(y.sum()), float(y.numel())), dtype=torch.float64)


def global_pos_weight(labels: torch.Tensor, group=None) -> torch.Tensor:
    """`batch_pos_weight` over the GLOBAL batch (one 2-float all-reduce of the class counts), so DP-N
    optimises the same loss as single-GPU training on the concatenated batch."""
    y = labels.reshape(-1).float()
    cnt = torch.stack([y.sum(), torch.tensor(float(y.numel()), device=y.device)]).to(torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cnt, group=group)
    return torch.tensor(class_pos_weight(float(cnt[0]), float(cnt[1])), dtype=torch.float64)


class GradReducer:
    """Flat-arena gradient all-reduce with per-bucket overlap.

    Usage per step::

        reducer.prepare()                 # zero the arena and point every p.grad into it
        loss = model(x, y, w); loss.backward()
        reducer.finish()                  # wait for the in-flight buckets; grads are now the global mean
        optimizer.step()

    Buckets are ordered as backward produces them (head first, embeddings last)."""

    def __init__(self, model: torch.nn.Module, group=None, bucket_of: Optional[Callable[[str], str]] = None,
                 overlap: bool = True, arena=None):
        """`arena`: an `optim.FlatArena` (e.g. `FusedSGD(...).arena`) to share its flat gradient buffer;
        otherwise the reducer allocates its own.  Buckets are runs of consecutive parameters (registration
        order) with the same bucket key, so each bucket is one contiguous slice of the arena."""
        self.model = model
        self.group = group
        self.overlap = overlap
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        bucket_of = bucket_of or self.default_bucket
        total = sum(p.numel() for _, p in named)
        if arena is not None:
            if [id(p) for p in arena.params] != [id(p) for _, p in named]:
                raise ValueError("arena and model parameters differ")
            self.flat = arena.flat_grad
        else:
            self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.arena = arena
        self.views = {}
        self.ranges = {}
        self.bucket_names = []
        self._bucket_params = {}
        self._param_bucket = {}
        off = 0
        for n, p in named:
            key = bucket_of(n)
            if not self.bucket_names or not self.bucket_names[-1].startswith(key + "#"):
                self.bucket_names.append(f"{key}#{len(self.bucket_names)}")
                self.ranges[self.bucket_names[-1]] = (off, off)
                self._bucket_params[self.bucket_names[-1]] = []
            b = self.bucket_names[-1]
            self.views[n] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            self.ranges[b] = (self.ranges[b][0], off)
            self._bucket_params[b].append(p)
            self._param_bucket[id(p)] = b
        self._pending = {}
        self._seen = set()
        self._named = named
        self._handles: List = []
        self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for _, p in named]

    @staticmethod
    def default_bucket(name: str) -> str:
        parts = name.split(".")
        if "layer" in parts:
            i = parts.index("layer")
            return ".".join(parts[:i + 2])          # one bucket per encoder Block
        if "embeddings" in parts:
            return "embeddings"
        return "head+norm"

    def prepare(self, defer: bool = False):
        """Start a step: zero the arena, point every p.grad into it.  With bucket overlap EXACTLY ONE backward pass may
        run between prepare() and finish(): a bucket is all-reduced as soon as its last gradient has been reported, so
        a second backward (gradient accumulation, the reference's --gradient_accumulation_steps) would add into
        gradients that are already reduced or in flight.  For accumulation pass defer=True: nothing is launched until
        finish(), which then all-reduces the whole arena once."""
        self.flat.zero_()
        for n, p in self._named:
            p.grad = self.views[n]
        self._pending = {b: len(ps) for b, ps in self._bucket_params.items()}
        self._handles = []
        self._seen = set()
        self._defer = bool(defer)
        self._launched = set()
        # parameters whose gradient the kernels accumulate in place never reach autograd's hooks:
        # functional._grad_done reports them here instead
        from . import functional as F
        F._STATE["grad_hook"] = self._on_grad

    def _on_grad(self, p):
        # a parameter can be reported twice per step: by the kernels' in-place accumulation path
        # (functional._grad_done) and by autograd's post-accumulate hook, which also fires for a None gradient
        if id(p) in self._seen:
            return
        self._seen.add(id(p))
        b = self._param_bucket.get(id(p))
        if b is None:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.overlap and not getattr(self, "_defer", False):
            self._launch(b)

    def _launch(self, b):
        _, ws = world()
        self._launched.add(b)
        if ws <= 1:
            return
        s, e = self.ranges[b]
        chunk = self.flat[s:e]
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                h = dist.all_reduce(chunk, group=self.group, async_op=True)
        else:
            h = dist.all_reduce(chunk, group=self.group, async_op=True)
        self._handles.append(h)

    def finish(self, scale: bool = True):
        """Waits for every bucket.  scale=True divides by world_size here (plain torch optimizers);
        scale=False leaves the SUM in the arena for a fused optimizer's `grad_scale=1/world_size`."""
        _, ws = world()
        if ws <= 1:
            return
        if not self.overlap or getattr(self, "_defer", False):
            self._handles.append(dist.all_reduce(self.flat, group=self.group, async_op=True))
        else:
            for b, left in self._pending.items():
                if left != 0:          # parameters without a gradient this step (unused): reduce anyway
                    self._launch(b)
        for h in self._handles:
            h.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        if scale:
            self.flat.mul_(1.0 / ws)
        self._handles = []

    def remove(self):
        for h in self._hooks:
            h.remove()
        from . import functional as F
        if F._STATE.get("grad_hook") == self._on_grad:
            F._STATE["grad_hook"] = None


# ----------------------------------------------------------------------------- ensemble sharding
def partition_work(costs: Sequence[float], batch: int, parts: int) -> List[List[Tuple[int, int, int]]]:
    """Cut the (member, volume) work list - member j's `batch` volumes cost `costs[j]` each, members laid
    end to end - into `parts` contiguous chunks of (nearly) equal cost.  Returns, per part, a list of
    (member, b0, b1) slices.  Every (member, volume) pair appears in exactly one part."""
    total = float(sum(costs)) * batch
    bounds = [total * r / parts for r in range(parts + 1)]
    out: List[List[Tuple[int, int, int]]] = [[] for _ in range(parts)]
    # position of pair (j, b) on the cost axis: sum_{i<j} costs[i]*batch + b*costs[j]
    cuts = []   # integer pair index boundaries
    flat_prefix = [0.0]
    for c in costs:
        flat_prefix.append(flat_prefix[-1] + c * batch)
    for r in range(parts + 1):
        t = bounds[r]
        j = 0
        while j < len(costs) - 1 and flat_prefix[j + 1] <= t:
            j += 1
        b = int(round((t - flat_prefix[j]) / costs[j])) if costs[j] > 0 else 0
        b = max(0, min(batch, b))
        cuts.append(j * batch + b)
    cuts[0], cuts[-1] = 0, len(costs) * batch
    for r in range(parts):
        lo, hi = cuts[r], max(cuts[r], cuts[r + 1])
        while lo < hi:
            j = lo // batch
            b0 = lo - j * batch
            b1 = min(batch, b0 + (hi - lo))
            out[r].append((j, b0, b1))
            lo += b1 - b0
    return out


def partition_batch_major(n_members: int, batch: int, parts: int) -> List[List[Tuple[int, int, int]]]:
    """The (member, batch-slice) work list cut BATCH-major: part r runs every member on volumes
    [r*batch/parts, (r+1)*batch/parts).  Every part costs the same whatever the members' FLOPs (3 members do not have
    to divide 2/4/8 ranks), and a rank needs only its own 1/parts of the input batch - with the member-major cut of
    `partition_work` a rank that holds one whole member needs the whole batch."""
    out: List[List[Tuple[int, int, int]]] = []
    for r in range(parts):
        b0, b1 = batch * r // parts, batch * (r + 1) // parts
        out.append([(j, b0, b1) for j in range(n_members)] if b1 > b0 else [])
    return out


class ShardedEnsemble:
    """TransformerEnsemble.forward (modeling.py:353-356) with members x batch-slices spread over the ranks.

    `ensemble` is a `TransformerEnsemble` whose members all live on this rank's device (weights are small:
    <= 60 MB per member).  Every rank is handed the same input batch `x`:
      * a DEVICE tensor: used as is;
      * a HOST tensor (pinned for speed; fp32 or uint8): only the batch slices of this rank's own
        (member, slice) chunk are copied to the device - with N ranks each ships ~3/N of a batch, not all of it.
    The member forwards of one rank run CONCURRENTLY on side streams when their slices are too small to fill the
    GPU on their own (the reference runs them back to back, modeling.py:354)."""

    CONCURRENT_MAX_SLICE = 256          # volumes: 256 x 65 rows = 130 row tiles < 148 SMs

    def __init__(self, ensemble, costs: Optional[Sequence[float]] = None, group=None, graphs: bool = True,
                 concurrent: Optional[bool] = None, partition: str = "batch"):
        """graphs=True replays each member's forward from a CUDA graph (one per member and slice shape): a rank's
        share of a batch is small, so launch overhead would otherwise dominate.  concurrent: None = automatic.
        partition: "batch" (default: `partition_batch_major`) or "member" (`partition_work`, FLOP-balanced contiguous
        chunks of the member-major list)."""
        self.ensemble = ensemble
        self.group = group
        self.partition = partition
        self._copy_stream = None
        self._stage_sets = {}
        self._stage_turn = 0
        m = len(ensemble.transformers)
        self.costs = list(costs) if costs is not None else [1.0] * m
        self._graphed = None
        self.concurrent = concurrent
        self._plans = {}
        self._streams = None
        params = list(ensemble.parameters()) if hasattr(ensemble, "parameters") else []
        self._cuda = bool(params) and params[0].is_cuda
        self._graphed_all = None
        if graphs and self._cuda:
            from .graphs import GraphedInference, GraphedMembers
            self._graphed = [GraphedInference(t) for t in ensemble.transformers]
            # batch-major sharding: every member runs on the same slice -> ONE graph with the members as parallel
            # branches (concurrent=False keeps the members back to back, for A/B timing)
            self._graphed_all = {c: GraphedMembers(ensemble, concurrent=c) for c in (True, False)}
        self.graphed = self._graphed is not None

    def graph_runners(self):
        if self._graphed is None:
            return []
        return list(self._graphed) + list(self._graphed_all.values())

    # ---- static plan per (batch size, world size)
    def _plan(self, B: int, device):
        rank, ws = world()
        key = (B, ws, rank, str(device))
        pl = self._plans.get(key)
        if pl is not None:
            return pl
        m = len(self.ensemble.transformers)
        parts = self._parts(B, ws)
        maxlen = max(1, max(sum(b1 - b0 for _, b0, b1 in p) for p in parts))
        mine = [(j, b0, b1) for j, b0, b1 in parts[rank] if b1 > b0]
        # merged batch intervals this rank needs (slices of consecutive members overlap or touch)
        iv = sorted((b0, b1) for _, b0, b1 in mine)
        merged = []
        for lo, hi in iv:
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        # out[b, j] = gathered[perm[b*m + j]]
        perm = torch.empty(B * m, dtype=torch.int64)
        for r, p in enumerate(parts):
            off = r * maxlen
            for j, b0, b1 in p:
                if b1 > b0:
                    perm[torch.arange(b0, b1) * m + j] = torch.arange(off, off + (b1 - b0))
                    off += b1 - b0
        pl = dict(parts=parts, maxlen=maxlen, mine=mine, merged=[tuple(v) for v in merged], perm=perm.to(device))
        self._plans[key] = pl
        return pl

    def _parts(self, B: int, ws: int):
        if self.partition == "member":
            return partition_work(self.costs, B, ws)
        return partition_batch_major(len(self.ensemble.transformers), B, ws)

    def gather_bytes(self, B: int) -> int:
        _, ws = world()
        parts = self._parts(B, ws)
        return 4 * ws * max(1, max(sum(b1 - b0 for _, b0, b1 in p) for p in parts))

    def h2d_bytes(self, B: int, bytes_per_volume: int = 327680) -> int:
        """Host-to-device bytes THIS rank ships for one batch handed in as a host tensor."""
        dev = next(self.ensemble.parameters()).device
        return sum(hi - lo for lo, hi in self._plan(B, dev)["merged"]) * bytes_per_volume

    class Staged:
        """A host batch on its way to the device (see `stage`): pass it to the call instead of the tensor."""

        def __init__(self, B, views, event, key):
            self.B, self.views, self.event, self.key = B, views, event, key
            self.shape = (B,)

    def stage(self, x: torch.Tensor) -> "ShardedEnsemble.Staged":
        """Starts the host-to-device copy of THIS rank's slices of the host batch `x` on a copy stream and returns a
        handle; `self(handle)` waits for it.  Two staging-buffer sets alternate, so the copy of batch s+1 can overlap
        the member forwards of batch s (the caller keeps at most one batch in flight ahead of the one it computes)."""
        dev = next(self.ensemble.parameters()).device
        B = x.shape[0]
        pl = self._plan(B, dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        turn = self._stage_turn
        self._stage_turn ^= 1
        key = (turn, B, x.dtype, tuple(x.shape[1:]))
        ent = self._stage_sets.get(key)
        if ent is None:
            bufs = [(lo, hi, torch.empty((hi - lo,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)) for lo, hi in pl["merged"]]
            ent = dict(bufs=bufs, free=torch.cuda.Event())
            ent["free"].record(torch.cuda.current_stream(dev))
            self._stage_sets[key] = ent
        cs = self._copy_stream
        cs.wait_event(ent["free"])               # the forwards that read this set two batches ago are done
        with torch.cuda.stream(cs):
            for lo, hi, buf in ent["bufs"]:
                buf.copy_(x[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        views = {}
        for _, b0, b1 in pl["mine"]:
            for lo, hi, buf in ent["bufs"]:
                if lo <= b0 and b1 <= hi:
                    views[(b0, b1)] = buf[b0 - lo:b1 - lo]
                    break
        return ShardedEnsemble.Staged(B, views, ev, key)

    @torch.no_grad()
    def member_logits(self, x: torch.Tensor) -> torch.Tensor:
        rank, ws = world()
        dev = next(self.ensemble.parameters()).device if hasattr(self.ensemble, "parameters") else x.device
        handle = None
        if isinstance(x, ShardedEnsemble.Staged):
            handle = x
        elif x.device != dev and dev.type == "cuda":
            handle = self.stage(x)               # host batch: ship this rank's slices now
        B = handle.B if handle is not None else x.shape[0]
        pl = self._plan(B, dev)
        mine = pl["mine"]
        staged = None
        if handle is not None:
            torch.cuda.current_stream(dev).wait_event(handle.event)
            staged = handle.views
        conc = self.concurrent
        m_all = len(self.ensemble.transformers)
        same_slice = len(mine) == m_all and len({(b0, b1) for _, b0, b1 in mine}) == 1 and [j for j, _, _ in mine] == list(range(m_all))
        if self._graphed_all is not None and same_slice:
            # one graph launch: all members of this rank's slice as parallel branches
            _, b0, b1 = mine[0]
            if conc is None:
                conc = (b1 - b0) <= self.CONCURRENT_MAX_SLICE
            xs = staged[(b0, b1)] if staged is not None else x[b0:b1]
            local = self._graphed_all[bool(conc)](xs)                    # (n, m)
            buf = torch.zeros(pl["maxlen"], device=dev, dtype=torch.float32)
            buf[:local.numel()] = local.t().reshape(-1)                  # member-major inside the rank's chunk (perm layout)
            if handle is not None:
                self._stage_sets[handle.key]["free"].record(torch.cuda.current_stream(dev))
            return self._gather(buf, pl, B, ws, dev)
        buf = torch.zeros(pl["maxlen"], device=dev, dtype=torch.float32)
        if conc is None:
            conc = self._cuda and len(mine) > 1 and max(b1 - b0 for _, b0, b1 in mine) <= self.CONCURRENT_MAX_SLICE
        if conc and self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(len(self.ensemble.transformers))]
        main = torch.cuda.current_stream(dev) if self._cuda else None
        off = 0
        used = []
        for k, (j, b0, b1) in enumerate(mine):
            member = self._graphed[j] if self._graphed is not None else self.ensemble.transformers[j]
            xs = staged[(b0, b1)] if staged is not None else x[b0:b1]
            if conc:
                s = self._streams[k % len(self._streams)]
                if s not in used:
                    s.wait_stream(main)          # fork: every side stream starts from the main stream's state
                    used.append(s)
                with torch.cuda.stream(s):
                    buf[off:off + (b1 - b0)] = member(xs)[0].reshape(-1).float()
            else:
                buf[off:off + (b1 - b0)] = member(xs)[0].reshape(-1).float()
            off += b1 - b0
        for s in used:                           # join only after ALL members were launched
            main.wait_stream(s)
        if handle is not None:
            self._stage_sets[handle.key]["free"].record(main)      # this staging set may be overwritten again
        return self._gather(buf, pl, B, ws, dev)

    def _gather(self, buf, pl, B, ws, dev):
        if ws > 1:
            gathered = torch.empty(ws * buf.numel(), device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(gathered, buf, group=self.group)
        else:
            gathered = buf
        return gathered[pl["perm"]].view(B, len(self.ensemble.transformers))

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        from . import functional as F
        feats = self.member_logits(x)
        return F.MetaFn.apply(feats, self.ensemble.classifier.weight, self.ensemble.classifier.bias)

    @property
    def copy_stream(self):
        return self._copy_stream


# ----------------------------------------------------------------------------- sweep packing
def pack_jobs(costs: Sequence[float], n_workers: int) -> List[List[int]]:
    """Longest-processing-time-first packing of independent jobs (the 18 configs x folds of
    train_baseline_cv.py, or bootstrap replicas) onto workers; returns job indices per worker."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * n_workers
    out: List[List[int]] = [[] for _ in range(n_workers)]
    for i in order:
        w = min(range(n_workers), key=lambda k: load[k])
        out[w].append(i)
        load[w] += costs[i]
    return out

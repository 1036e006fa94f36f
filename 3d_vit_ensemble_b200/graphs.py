"""CUDA-graph execution of the path (launch-bound regimes: the reference trains at batch 4 and validates
at batch 1, train_baseline_cv.py:240, :64-101).  Every kernel of libvit3d_sm100.so is launched on the
caller's stream without synchronising, so a whole forward - or a whole training step including the
optimizer - is capturable; values that change from step to step (dropout step, per-batch class weight,
learning rate, Adam step) are read from device scalars instead of being baked into the graph."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from . import functional as F
from . import fused_train


class GraphedInference:
    """`model(x)` (eval, no_grad) replayed from a CUDA graph, one graph per input shape.

        g = GraphedInference(model)
        logits, attn, enc = g(x)            # outputs are static buffers, overwritten by the next call
    """

    def __init__(self, model: torch.nn.Module, warmup: int = 2):
        self.model = model
        self.warmup = warmup
        self._graphs = {}
        self._params = list(model.parameters())
        self._sig = None
        self.launches_per_replay = 0
        self.replays = 0
        self.recaptures = 0

    def _weights_signature(self):
        """The captured kernels read cached low-precision weight shadows (functional._LP_CACHE / _QKV_CACHE); an
        optimizer step, load_state_dict, .to() or set_precision makes the eager path rebuild those caches and free
        the old tensors, so a graph captured before would compute with stale weights - or freed memory."""
        return (tuple(p._version for p in self._params), tuple(p.data_ptr() for p in self._params),
                F._STATE.get("epoch", 0), getattr(self.model, "precision", None))

    @torch.no_grad()
    def __call__(self, x: torch.Tensor):
        sig = self._weights_signature()
        if sig != self._sig:
            if self._graphs:
                self._graphs = {}          # weights changed since capture: every graph is stale, capture again
                self.recaptures += 1
            self._sig = sig
        key = (tuple(x.shape), x.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            self.model.eval()
            static_x = x.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(self.warmup):          # populates the weight-shadow cache: no casts in the graph
                    self.model(static_x)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.lib().vit3d_launch_count()
            with torch.cuda.graph(g):
                out = self.model(static_x)
            self.launches_per_replay = _lib.lib().vit3d_launch_count() - n0     # kernels recorded in the graph
            ent = (g, static_x, out)
            self._graphs[key] = ent
        g, static_x, out = ent
        if x.data_ptr() != static_x.data_ptr():       # a caller that fills `input_like(x)` in place skips this copy
            static_x.copy_(x, non_blocking=True)
        g.replay()
        self.replays += 1
        return out

    def input_like(self, x: torch.Tensor) -> torch.Tensor:
        """The graph's own input buffer for inputs shaped like `x` (captures the graph on first use, with x's
        contents).  Writing the next batch straight into it (e.g. as the target of the host-to-device copy) and
        passing it to the call saves the device-to-device copy of the batch - 0.67 GB of HBM traffic, ~0.1 ms
        of a 1.8 ms step at batch 1024."""
        key = (tuple(x.shape), x.dtype)
        if key not in self._graphs:
            self(x)
        return self._graphs[key][1]


class GraphedMembers:
    """The member forwards of a TransformerEnsemble (modeling.py:354, `[t(x)[0] for t in transformers]`) replayed from
    ONE CUDA graph in which every member is a parallel branch (forked streams inside the capture): one graph launch
    per call, and the members overlap on the GPU when a batch is too small to fill it with one member.

        g = GraphedMembers(ensemble)
        logits = g(x)            # (B, n_members) fp32, static buffer
    """

    def __init__(self, ensemble, warmup: int = 2, concurrent: bool = True):
        self.ensemble = ensemble
        self.members = list(ensemble.transformers)
        self.warmup = warmup
        self.concurrent = concurrent
        self._graphs = {}
        self._params = [p for m in self.members for p in m.parameters()]
        self._sig = None
        self.launches_per_replay = 0
        self.replays = 0
        self.recaptures = 0

    def _weights_signature(self):
        return (tuple(p._version for p in self._params), tuple(p.data_ptr() for p in self._params),
                F._STATE.get("epoch", 0), tuple(getattr(m, "precision", None) for m in self.members))

    def _run(self, x, streams):
        if x.dtype == torch.uint8:                  # N2: convert once for all members
            x = F.u8_volumes_to_f32(x, self.members[0].input_mean)
        main = torch.cuda.current_stream()
        outs = []
        if streams is None:
            outs = [m(x)[0] for m in self.members]
        else:
            for m, s in zip(self.members, streams):
                s.wait_stream(main)
                with torch.cuda.stream(s):
                    outs.append(m(x)[0])
            for s in streams:
                main.wait_stream(s)
        return torch.cat([o.reshape(x.shape[0], -1).float() for o in outs], dim=1)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor):
        sig = self._weights_signature()
        if sig != self._sig:
            if self._graphs:
                self._graphs = {}
                self.recaptures += 1
            self._sig = sig
        key = (tuple(x.shape), x.dtype)
        ent = self._graphs.get(key)
        if ent is None:
            for m in self.members:
                m.eval()
            static_x = x.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self._run(static_x, None)
            torch.cuda.current_stream().wait_stream(s)
            streams = [torch.cuda.Stream() for _ in self.members] if (self.concurrent and len(self.members) > 1) else None
            g = torch.cuda.CUDAGraph()
            n0 = _lib.lib().vit3d_launch_count()
            with torch.cuda.graph(g):
                out = self._run(static_x, streams)
            self.launches_per_replay = _lib.lib().vit3d_launch_count() - n0
            ent = (g, static_x, out)
            self._graphs[key] = ent
        g, static_x, out = ent
        if x.data_ptr() != static_x.data_ptr():
            static_x.copy_(x, non_blocking=True)
        g.replay()
        self.replays += 1
        return out


class GraphedTrainStep:
    """One captured training step: loss = model(x, y, pos_weight); loss.backward(); optimizer.step().

    `optimizer` must be a `optim.FusedSGD` / `FusedAdam` (flat arena: the gradient zeroing, the step and
    the weight-shadow refresh are all inside the graph).  The warm-up iterations before capture ARE
    training steps on the first batch passed in.

        step = GraphedTrainStep(model, opt)
        loss = step(x, y, pos_weight)       # pos_weight: float / 0-dim tensor / None; loss: static 0-dim tensor
        step.set_lr(lr)                     # follow an LR schedule without re-capturing
    """

    def __init__(self, model: torch.nn.Module, optimizer, warmup: int = 3, grad_scale: float = 1.0,
                 data_parallel: bool = False, group=None, overlap: bool = False):
        """data_parallel=True (torch.distributed initialised): the graph holds zero_grad + forward + backward;
        each replay is followed by the all-reduce of the flat gradient arena (60 MB at conf 18: ~0.3 ms over NVLink
        against a 3.4 ms step) and the fused optimizer step with grad_scale = 1/world_size.  overlap=True (fused BF16
        step, >= 4 Blocks): the step is captured as two graphs cut in the middle of the backward and the upper half of
        the arena is all-reduced while the second graph runs - equivalent (tools/check_multi_gpu.py) and measured
        EQUAL at N = 2 (3.61 ms either way, tools/dp_overlap_timing.py): the backward's kernels fill every SM's shared
        memory, NCCL's CTAs only get in between kernels; off by default.  Pass the GLOBAL-batch pos_weight
        (dist.global_pos_weight)."""
        self.model = model
        self.opt = optimizer
        self.warmup = max(1, warmup)
        self.grad_scale = grad_scale
        self.dp = bool(data_parallel)
        self.group = group
        self.overlap = bool(overlap)
        self._graph = None
        self._graphs = None
        self.launches_per_replay = 0
        self.replays = 0
        dev = optimizer.arena.flat.device
        self.lr_dev = torch.tensor([float(optimizer.param_groups[0]["lr"])], device=dev, dtype=torch.float32)
        self.pw_dev = torch.ones(1, device=dev, dtype=torch.float32)
        # dropout mask offset and Adam's t: starts at the optimizer's step count, so a resumed run (or eager steps
        # taken before the capture) continues Adam's bias correction and the mask sequence instead of restarting them
        self.step_dev = torch.full((1,), int(getattr(optimizer, "_steps", 0)), device=dev, dtype=torch.int32)
        self.fused = False
        # per-step scalars (class weight, learning rate) reach the device through a small ring of PINNED host slots:
        # a copy from pageable memory would make the host wait for the stream (it serialises concurrent job streams),
        # a fill kernel would be a framework kernel in the step
        self._scalars = torch.empty(128, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.empty(128)
        self._scalar_i = 0
        optimizer.lr_dev = self.lr_dev
        if hasattr(optimizer, "exp_avg"):
            optimizer.step_dev = self.step_dev

    def _put_scalar(self, dst: torch.Tensor, value: float):
        i = self._scalar_i
        self._scalar_i = (i + 1) % self._scalars.numel()
        self._scalars[i] = float(value)
        dst.copy_(self._scalars[i:i + 1], non_blocking=True)

    def set_lr(self, lr: float):
        self._put_scalar(self.lr_dev, lr)
        self.opt.param_groups[0]["lr"] = float(lr)

    def _fwd_bwd(self):
        pw = self.pw_dev if self.use_pw else None
        if self.fused:
            # fused BF16 step: shadow refresh (+ step counter), memset of the gradient arena, forward and backward
            # as explicit kernel sequences - no autograd, no framework kernels inside the graph
            fused_train.plan_of(self.model).refresh(self.step_dev, force=True)
            self.opt.zero_grad()
            return fused_train.loss_and_grads(self.model, self.x, self.y, pw)
        self.step_dev.add_(1)
        self.opt.zero_grad()
        loss = self.model(self.x, self.y, pw)
        loss.backward()
        return loss

    def _finish(self):
        """Outside the graph in data-parallel mode: gradient all-reduce + optimizer step."""
        import torch.distributed as dist
        ws = dist.get_world_size(self.group)
        dist.all_reduce(self.opt.arena.flat_grad, group=self.group)
        self.opt.step(grad_scale=self.grad_scale / ws)

    # ---- data-parallel overlap: the step is captured as TWO graphs cut inside the backward; the gradients of the
    # upper half of the model (head, encoder_norm, Blocks >= L/2 - the tail of the flat arena, parameters are laid out
    # in model order) are all-reduced on NCCL's stream while the second graph computes the lower half
    def _dp_cut(self):
        if not (self.dp and self.fused) or F._STATE.get("wgrad_partial") or not self.overlap:
            return None
        layers = list(self.model.transformer.encoder.layer)
        if len(layers) < 4:
            return None
        cut = len(layers) // 2
        first = next(layers[cut].parameters())
        off = 0
        for p_ in self.opt.arena.params:
            if p_ is first:
                return cut, off
            off += p_.numel()
        return None

    def _capture_segments(self, cut):
        pw = self.pw_dev if self.use_pw else None
        pool = torch.cuda.graph_pool_handle()
        gen = fused_train.loss_and_grads_steps(self.model, self.x, self.y, pw, (cut,))
        self._graphs = [torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()]
        with torch.cuda.graph(self._graphs[0], pool=pool):
            fused_train.plan_of(self.model).refresh(self.step_dev, force=True)
            self.opt.zero_grad()
            self.loss = next(gen)
        with torch.cuda.graph(self._graphs[1], pool=pool):
            try:
                next(gen)
                raise RuntimeError("backward yielded more segments than graphs")
            except StopIteration:
                pass

    def _replay_segments(self):
        import torch.distributed as dist
        ws = dist.get_world_size(self.group)
        fg = self.opt.arena.flat_grad
        self._graphs[0].replay()
        w0 = dist.all_reduce(fg[self._cut_off:], group=self.group, async_op=True)
        self._graphs[1].replay()
        w1 = dist.all_reduce(fg[:self._cut_off], group=self.group, async_op=True)
        w0.wait()
        w1.wait()
        self.opt.step(grad_scale=self.grad_scale / ws)

    def _one(self):
        loss = self._fwd_bwd()
        if self.dp:
            self._finish()
        else:
            self.opt.step(grad_scale=self.grad_scale)
        return loss

    def __call__(self, x, y, pos_weight=None):
        if self._graph is None:
            self.model.train()
            self.x, self.y = x.clone(), y.clone().float()
            self.use_pw = pos_weight is not None
            if self.use_pw:
                self.pw_dev.fill_(float(pos_weight))
            F._STATE["step_dev"] = self.step_dev
            self.fused = fused_train.supported(self.model, self.x)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self._one()
            torch.cuda.current_stream().wait_stream(s)
            F.invalidate_weight_shadows()       # the captured step must re-derive the bf16 shadows itself
            self._graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib().vit3d_launch_count()
            cut = self._dp_cut()
            if cut is not None:
                self._cut_off = cut[1]
                self._capture_segments(cut[0])
            else:
                self._graphs = None
                with torch.cuda.graph(self._graph):
                    self.loss = self._fwd_bwd() if self.dp else self._one()
            self.launches_per_replay = _lib.lib().vit3d_launch_count() - n0
            if not self.dp:
                self.opt._steps -= 1            # the capture recorded the optimizer launch, it did not run it
            F.invalidate_weight_shadows()       # shadows made during capture live in the graph's pool
            self._replay()                      # capture records, it does not execute: run the step now
            return self.loss
        if (pos_weight is not None) != self.use_pw:
            raise ValueError("GraphedTrainStep was captured %s pos_weight" % ("with" if self.use_pw else "without"))
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        if self.use_pw:
            self._put_scalar(self.pw_dev, float(pos_weight))
        self._replay()
        return self.loss

    def _replay(self):
        if self._graphs is not None:
            self.replays += 1
            self._replay_segments()             # two graphs, the all-reduces of the two arena halves between / after them
            F.invalidate_weight_shadows()
            return
        self._graph.replay()
        self._after_replay()

    def _after_replay(self):
        self.replays += 1
        if self.dp:
            self._finish()                      # counts the optimizer step itself
        else:
            self.opt._steps += 1                # the captured optimizer launch ran again
        F.invalidate_weight_shadows()           # weights moved behind torch's version counters (eager users re-derive)

    def input_buffers(self):
        """(x, y) the graph reads: write the next batch straight into them (e.g. as host-to-device copy targets)
        and pass them to the call to skip the device-to-device copies."""
        return self.x, self.y

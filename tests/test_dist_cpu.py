"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: flat-arena gradient all-reduce with
bucket overlap, global class weight, ensemble work partition + all-gather, sweep packing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vit3d_b200
from vit3d_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


class Toy(torch.nn.Module):
    """Parameter names shaped like the ViT's so the default bucketing (per encoder Block) is exercised."""

    def __init__(self):
        super().__init__()
        self.transformer = torch.nn.Module()
        self.transformer.embeddings = torch.nn.Linear(6, 8)
        self.transformer.encoder = torch.nn.Module()
        self.transformer.encoder.layer = torch.nn.ModuleList([torch.nn.Linear(8, 8) for _ in range(3)])
        self.head = torch.nn.Linear(8, 1)

    def forward(self, x):
        h = self.transformer.embeddings(x)
        for l in self.transformer.encoder.layer:
            h = torch.tanh(l(h)) + h
        return self.head(h)


def _dp_equals_big_batch(rank, world):
    torch.manual_seed(0)
    model = Toy()
    x = torch.randn(8, 6)
    y = (torch.arange(8) % 3 == 0).float()
    # reference: single process, whole batch, global class weight
    ref = Toy()
    ref.load_state_dict(model.state_dict())
    pw = y.numel() / (2.0 * float(y.sum()))          # sklearn 'balanced' weights[1] (train_baseline_cv.py:168-169)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(ref(x).reshape(-1), y, pos_weight=torch.tensor(pw))
    loss.backward()
    # DP: each rank sees its shard
    xs, ys = x[rank::world], y[rank::world]
    w = D.global_pos_weight(ys)
    assert abs(float(w) - pw) < 1e-12
    for overlap in (True, False):
        red = D.GradReducer(model, overlap=overlap)
        assert red.bucket_names[0].startswith("transformer.embeddings") or red.bucket_names[0].startswith("embeddings")
        assert len(red.bucket_names) == 5
        red.prepare()
        l = torch.nn.functional.binary_cross_entropy_with_logits(model(xs).reshape(-1), ys, pos_weight=w.float())
        l.backward()
        red.finish()
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            assert p.grad.data_ptr() == red.views[n].data_ptr()          # gradients live in the flat arena
            torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-6)
        red.remove()


def test_grad_reducer_matches_single_process_big_batch():
    spawn(_dp_equals_big_batch)


def _sharded_ensemble(rank, world):
    class Member(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k = k

        def forward(self, x):
            return (x.reshape(x.shape[0], -1).sum(1, keepdim=True) * self.k, [], None)

    class Ens:
        transformers = [Member(1.0), Member(-2.0), Member(0.5)]

    x = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4)
    want = torch.cat([m(x)[0] for m in Ens.transformers], dim=1)
    for mode in ("batch", "member"):
        se = D.ShardedEnsemble(Ens(), costs=[1.09, 1.44, 1.01], partition=mode)
        torch.testing.assert_close(se.member_logits(x), want)


def test_sharded_ensemble_allgather():
    spawn(_sharded_ensemble)


@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("batch", [1, 3, 64])
def test_partition_covers_every_pair_once_and_balances(parts, batch):
    costs = [1.0903, 1.4397, 1.0135]
    p = D.partition_work(costs, batch, parts)
    seen = set()
    loads = []
    for chunk in p:
        load = 0.0
        for j, b0, b1 in chunk:
            assert 0 <= b0 <= b1 <= batch
            for b in range(b0, b1):
                assert (j, b) not in seen
                seen.add((j, b))
            load += costs[j] * (b1 - b0)
        loads.append(load)
    assert len(seen) == 3 * batch
    if batch >= 64:
        assert max(loads) <= 1.05 * sum(loads) / parts + max(costs)


@pytest.mark.parametrize("parts", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("batch", [1, 3, 64, 512])
def test_batch_major_partition_covers_every_pair_once_and_ships_one_slice_per_rank(parts, batch):
    p = D.partition_batch_major(3, batch, parts)
    seen = set()
    for chunk in p:
        slices = {(b0, b1) for _, b0, b1 in chunk}
        assert len(slices) <= 1                          # a rank needs ONE contiguous slice of the input batch
        for j, b0, b1 in chunk:
            for b in range(b0, b1):
                assert (j, b) not in seen
                seen.add((j, b))
    assert len(seen) == 3 * batch
    sizes = [sum(b1 - b0 for _, b0, b1 in c) for c in p]
    assert max(sizes) - min(sizes) <= 3


def test_pack_jobs():
    costs = [5.95] * 5 + [3.27] * 5 + [2.2] * 8
    packs = D.pack_jobs(costs, 8)
    assert sorted(i for p in packs for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in packs]
    assert max(loads) - min(loads) <= max(costs)


def test_grad_reducer_counts_each_parameter_once():
    """A parameter may be reported by the in-place accumulation path AND by autograd's hook (which also
    fires for a None gradient): the bucket countdown must not go negative (it would all-reduce twice)."""
    m = Toy()
    red = D.GradReducer(m)
    red.prepare()
    for p in m.parameters():
        red._on_grad(p)
        red._on_grad(p)
    assert all(v == 0 for v in red._pending.values()), red._pending
    red.remove()


def test_pos_weight_matches_sklearn_balanced():
    """dist.batch_pos_weight == what train_baseline_cv.py:168-169 computes with sklearn, including the
    single-class batches (weights has one entry, 1.0) - ADVICE r1."""
    import numpy as np
    from sklearn.utils import class_weight
    from oracle import vit3d_oracle as O
    for labels in ([0, 0, 0, 1], [1, 1, 0, 1], [0, 1], [1, 1, 1, 1], [0, 0, 0, 0], [0, 1, 1, 0, 1, 0, 0, 0]):
        y = np.asarray(labels, dtype=np.float32)
        w = class_weight.compute_class_weight(class_weight="balanced", classes=np.unique(y), y=y)
        want = float(w[1] if len(w) > 1 else w[0])
        assert abs(float(D.batch_pos_weight(torch.tensor(y))) - want) < 1e-12, labels
        assert abs(float(D.global_pos_weight(torch.tensor(y))) - want) < 1e-12, labels
        assert abs(float(O.sklearn_pos_weight(torch.tensor(y))) - want) < 1e-12, labels


def _accumulate_two_backwards(rank, world):
    """gradient accumulation (two backward passes per step) with prepare(defer=True) equals the big batch."""
    torch.manual_seed(0)
    model = Toy()
    ref = Toy()
    ref.load_state_dict(model.state_dict())
    x = torch.randn(8, 6)
    y = (torch.arange(8) % 3 == 0).float()
    loss = torch.nn.functional.binary_cross_entropy_with_logits(ref(x).reshape(-1), y)
    loss.backward()
    xs, ys = x[rank::world], y[rank::world]
    red = D.GradReducer(model, overlap=True)
    red.prepare(defer=True)
    half = xs.shape[0] // 2
    for a, b in ((0, half), (half, xs.shape[0])):          # two micro-batches, each scaled to its share of the mean
        l = torch.nn.functional.binary_cross_entropy_with_logits(model(xs[a:b]).reshape(-1), ys[a:b], reduction="sum") / xs.shape[0]
        l.backward()
    assert not red._launched                                # nothing was reduced before finish()
    red.finish()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    red.remove()


def test_grad_reducer_deferred_mode_supports_gradient_accumulation():
    spawn(_accumulate_two_backwards)

"""Host-side helpers of SURVEY §8(f): checkpoint interop / resume state (N4, CPU) and batched single-forward
validation (N3, GPU)."""
import os

import pytest
import torch

import vit3d_b200
from oracle import vit3d_oracle as O
from vit3d_b200 import workflow as W


def test_training_state_round_trip(tmp_path):
    """save_training_state / load_training_state restore weights, optimizer buffers, scheduler and step."""
    m = W.build_baseline(1)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=3, gamma=0.5)
    for p in m.parameters():
        p.grad = torch.full_like(p, 0.01)
    opt.step()
    sched.step()
    path = os.path.join(tmp_path, "state.pt")
    W.save_training_state(path, m, opt, sched, step=17, extra={"fold": 2})
    m2 = W.build_baseline(1)
    opt2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, step_size=3, gamma=0.5)
    assert W.load_training_state(path, m2, opt2, sched2) == 17
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert sched2.state_dict() == sched.state_dict()
    b1 = [s["momentum_buffer"] for s in opt.state.values()]
    b2 = [s["momentum_buffer"] for s in opt2.state.values()]
    assert all(torch.equal(x, y) for x, y in zip(b1, b2))
    assert not os.path.exists(path + ".tmp")


def test_ensemble_from_checkpoints_uses_modules(tmp_path):
    """Members restored from state_dict files are real modules with one logit each (the reference passes the
    return value of load_state_dict and in_features=3, train_ensemble_whole_dataset.py:50-52)."""
    paths, confs = [], [5, 11]
    for c in confs:
        sd = O.init_state_dict(vit3d_b200.north_star_config(c), seed=c)
        p = os.path.join(tmp_path, f"conf{c}.bin")
        torch.save(sd, p)
        paths.append(p)
    ens = W.ensemble_from_checkpoints(paths, confs, device="cpu")
    assert len(ens.transformers) == 2 and ens.classifier.in_features == 2
    sd5 = O.init_state_dict(vit3d_b200.north_star_config(5), seed=5)
    assert torch.equal(ens.transformers[0].head.weight, sd5["head.weight"])
    with pytest.raises(ValueError):
        W.ensemble_from_checkpoints(paths, [5], device="cpu")


@pytest.mark.gpu
def test_validate_matches_per_sample_loop():
    """validate(): batches + one forward per batch == the reference's per-sample loop (train_baseline_cv.py:64-101)."""
    cfg = vit3d_b200.get_config(16, 512, 2, 256, 8)
    m = vit3d_b200.models.modeling.VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16")
    m.load_state_dict(O.init_state_dict(cfg, seed=3))
    m.to("cuda:0")
    x = O.synth_volumes(11, seed=5)
    y = O.synth_labels(11)
    out = W.validate(m, x, y, batch_size=4)
    m.eval()
    with torch.no_grad():
        ref = torch.cat([m(x[i:i + 1].to("cuda:0"))[0].reshape(-1).cpu() for i in range(11)])
    assert out["logits"].shape == (11,) and out["features"].shape == (11, 256)
    assert float((out["logits"] - ref).abs().max()) <= 2e-2
    assert torch.equal(out["predicted"], (torch.sigmoid(out["logits"]) > 0.5).long())
    acc = float((out["predicted"] == y.long()).float().mean())
    assert abs(float(out["accuracy"]) - acc) < 1e-6
    assert 0.0 <= float(out["balanced_accuracy"]) <= 1.0

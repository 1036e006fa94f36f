"""Host-side helpers of SURVEY §8(f): checkpoint interop / resume state (N4, CPU) and batched single-forward
validation (N3, GPU)."""
import os

import pytest
import torch

import vit3d_b200
import vit3d_b200.models.modeling  # noqa: F401  (the reference-named module the workflow resolves lazily)
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


def _ref_checkpoint_fixture():
    import numpy as np
    from tests.helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "ref_ckpt_ensemble.npz"))
    paths = [os.path.join(GOLDEN, f"ref_ckpt_member{j}.bin") for j in range(3)]
    cfgs = [vit3d_b200.get_config(*[int(v) for v in a]) for a in g["cfg_args"]]
    return g, paths, cfgs


def test_reference_written_checkpoints_load_with_identical_keys():
    """tests/golden/ref_ckpt_member*.bin were written by the UNMODIFIED reference (oracle/gen_golden.py --checkpoints:
    `torch.save(model.state_dict(), path)`, train_baseline_cv.py:128-134): strict load, same keys / shapes / values."""
    g, paths, cfgs = _ref_checkpoint_fixture()
    ens = W.ensemble_from_checkpoints(paths, configs=cfgs, device="cpu", precision="fp32")
    for path, m in zip(paths, ens.transformers):
        sd = torch.load(path, map_location="cpu")
        mine = m.state_dict()
        assert list(sd.keys()) == list(mine.keys())
        for k in sd:
            assert sd[k].shape == mine[k].shape and torch.equal(sd[k], mine[k]), k


@pytest.mark.gpu
def test_reference_written_checkpoints_reproduce_reference_ensemble_outputs():
    """N4 end to end on the GPU: reference `.bin` files -> ensemble_from_checkpoints -> member logits and ensemble
    probabilities equal what the reference's own TransformerEnsemble computed from the same weights."""
    import numpy as np
    g, paths, cfgs = _ref_checkpoint_fixture()
    ens = W.ensemble_from_checkpoints(paths, configs=cfgs, device="cuda:0", precision="fp32")
    with torch.no_grad():
        ens.classifier.weight.copy_(torch.from_numpy(g["classifier_weight"]))
        ens.classifier.bias.copy_(torch.from_numpy(g["classifier_bias"]))
    ens.eval()
    x = O.synth_volumes(3, seed=5, kind="img").to("cuda:0")
    with torch.no_grad():
        members = torch.cat([t(x)[0] for t in ens.transformers], dim=1).cpu().numpy()
        out = ens(x).cpu().numpy()
    np.testing.assert_allclose(members, g["member_logits"], atol=1e-3)       # north_star: logits <= 1e-3 in fp32 mode
    np.testing.assert_allclose(out, g["out"], atol=1e-4)
    assert ((out > 0.5) == (g["out"] > 0.5)).all()

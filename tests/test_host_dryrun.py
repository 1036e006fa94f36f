"""Exercises the Python host logic end to end without a GPU (argument wiring of every C-ABI call,
autograd plumbing, shapes/dtypes).  See tests/dryrun.py."""
import pytest
import torch

import vit3d_b200
from tests.dryrun import dry_run
from vit3d_b200 import functional as F
from vit3d_b200.models.modeling import TransformerEnsemble, VisionTransformer


@pytest.mark.parametrize("prec", ["fp32", "bf16", "tf32"])
@pytest.mark.parametrize("args", [(16, 64, 2, 32, 4), (8, 48, 1, 8, 8)])
def test_train_and_eval_control_flow(prec, args):
    cfg = vit3d_b200.get_config(*args)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec)
    x = torch.randn(2, 1, 128, 128, 5)
    y = torch.tensor([0.0, 1.0])
    with dry_run() as calls:
        m.train()
        loss = m(x, y, torch.tensor(1.5, dtype=torch.float64))
        assert loss.dim() == 0
        loss.backward()
        for n, p in m.named_parameters():
            assert p.grad is not None and p.grad.shape == p.shape, n
        m.eval()
        with torch.no_grad():
            logits, attn, enc = m(x)
        P = (128 // args[0]) ** 2
        assert logits.shape == (2, 1) and enc.shape == (2, P + 1, args[3]) and len(attn) == args[2]
        assert attn[0].shape == (2, args[4], P + 1, P + 1)
    assert "vit3d_patch_embed_fwd" in calls and "vit3d_attn_bwd" in calls and "vit3d_dropout" in calls


def test_ensemble_control_flow():
    cfgs = [vit3d_b200.get_config(16, 64, 1, 32, 4), vit3d_b200.get_config(16, 32, 2, 32, 8)]
    ens = TransformerEnsemble(*[VisionTransformer(c, 128, zero_head=True, num_classes=1, precision="fp32") for c in cfgs],
                              in_features=1)
    x = torch.randn(3, 1, 128, 128, 5)
    with dry_run():
        out = ens(x)
        assert out.shape == (3, 1)
        out.sum().backward()
        assert ens.classifier.weight.grad.shape == (1, 2)

"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle (oracle/vit3d_oracle.py) and the committed golden vectors of the reference.

Tolerances (BASELINE.json north_star): patch ordering bit-exact; logits <= 1e-3 max-abs in
fp32/TF32 mode, <= 2e-2 in bf16; identical predicted classes."""
import numpy as np
import pytest
import torch

import vit3d_b200
from oracle import vit3d_oracle as O
from tests.helpers import CASES, assert_stats_close, case_setup, load_golden, stats, unpack_masks
from vit3d_b200 import functional as F
from vit3d_b200.models.modeling import (Attention, Block, Embeddings, Encoder, Mlp, TransformerEnsemble,
                                        VisionTransformer)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

LOGIT_TOL = {"fp32": 1e-3, "tf32": 1e-3, "bf16": 2e-2}


def build(name, prec, vis=True):
    cfg, sd, x, y, w = case_setup(name)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, vis=vis, precision=prec)
    m.load_state_dict(sd)
    m.to(DEV)
    return cfg, sd, m, x, y, w


# ----------------------------------------------------------------------------- a1
@pytest.mark.parametrize("patch", [(16, 16, 5), (8, 8, 5), (32, 16, 5)])
@pytest.mark.parametrize("B", [1, 3])
def test_patch_gather_bit_exact(patch, B):
    x = O.synth_volumes(B, seed=5, kind="unit")
    got = F.patch_gather(x.to(DEV), patch).cpu()
    assert torch.equal(got, O.patch_gather(x, patch))


def test_patch_gather_empty_batch():
    x = torch.zeros(0, 1, 128, 128, 5, device=DEV)
    assert F.patch_gather(x, (16, 16, 5)).shape == (0, 64, 1280)


# ----------------------------------------------------------------------------- whole model, fp32 exact path
@pytest.mark.parametrize("name", list(CASES))
def test_forward_fp32_vs_oracle_and_golden(name):
    g = load_golden(name)
    cfg, sd, m, x, y, w = build(name, "fp32")
    m.eval()
    with torch.no_grad():
        logits, probs, enc = m(x.to(DEV))
    lo, po, eo = O.vit_forward(sd, cfg, x)
    assert logits.shape == lo.shape and enc.shape == eo.shape and len(probs) == len(po)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=1e-3, rtol=0)      # reference golden
    np.testing.assert_allclose(logits.cpu().numpy(), lo.numpy(), atol=2e-4, rtol=0)       # oracle
    np.testing.assert_allclose(enc.cpu().numpy(), eo.numpy(), atol=5e-4, rtol=0)
    for a, b in zip(probs, po):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), atol=2e-5, rtol=0)
    assert ((torch.sigmoid(logits.cpu()) > 0.5) == (torch.sigmoid(lo) > 0.5)).all()
    xu = O.synth_volumes(x.shape[0], seed=43, kind="unit")
    with torch.no_grad():
        lu = m(xu.to(DEV))[0]
    np.testing.assert_allclose(lu.cpu().numpy(), g["logits_unit"], atol=1e-3, rtol=0)


def _grad_check(m, grads_oracle, rtol, what, atol_scale=1e-6):
    # atol_scale * (largest gradient norm): the key.bias gradients are exactly 0 in exact arithmetic
    gscale = max(float(v.norm()) for v in grads_oracle.values())
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        a, b = p.grad.detach().cpu(), grads_oracle[k]
        err = float((a - b).norm())
        ref = float(b.norm())
        assert err <= rtol * ref + atol_scale * gscale, (what, k, err, ref)


@pytest.mark.parametrize("name", ["tiny", "shipped", "shipped_p8", "conf5", "conf18"])
def test_loss_and_grads_fp32_eval_mode(name):
    g = load_golden(name)
    cfg, sd, m, x, y, w = build(name, "fp32")
    m.eval()
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w)
    assert abs(float(loss) - float(g["loss_eval"])) < 1e-4
    assert abs(float(loss) - float(lo)) < 2e-5
    _grad_check(m, go, 2e-3, name)
    for k, p in m.named_parameters():
        gs = g["grad_eval_stats/" + k]
        gscale = max(float(g["grad_eval_stats/" + kk][1]) for kk, _ in m.named_parameters())
        assert_stats_close(stats(p.grad), gs, 2e-3, 2e-6 * gscale, k)


@pytest.mark.parametrize("name", ["tiny", "conf5"])
def test_train_mode_with_reference_dropout_masks(name):
    """Dropout parity by mask injection: replay the masks the reference drew (golden fixture)."""
    g = load_golden(name)
    cfg, sd, m, x, y, w = build(name, "fp32")
    masks = unpack_masks(g)
    sites = {0: masks["emb"].reshape(-1)}
    for i in range(cfg.transformer["num_layers"]):
        sites[1 + 2 * i] = masks[("fc1", i)].reshape(-1)
        sites[2 + 2 * i] = masks[("fc2", i)].reshape(-1)
    m.train()
    with F.mask_injection(sites):
        loss = m(x.to(DEV), y.to(DEV), w)
        loss.backward()
    assert abs(float(loss) - float(g["loss_train"])) < 1e-4
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, masks=masks)
    assert abs(float(loss) - float(lo)) < 2e-5
    _grad_check(m, go, 2e-3, name)


def test_own_dropout_masks_are_consistent_and_calibrated():
    """The library's own Philox masks: keep-rate ~ 1-p, deterministic per (seed, site, step), and the
    forward/backward of a training step equals the oracle with those masks injected."""
    name = "tiny"
    cfg, sd, m, x, y, w = build(name, "fp32")
    m.train()
    torch.manual_seed(99)
    step0 = F._STATE["step"]
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    step = step0 + 1
    B, S, H, d = x.shape[0], 65, cfg.hidden_size, cfg.transformer["mlp_dim"]
    p = cfg.transformer["dropout_rate"]
    masks = {"emb": F.dropout_mask(B * S * H, p, 0, step, DEV).cpu().bool().reshape(B, S, H)}
    for i in range(cfg.transformer["num_layers"]):
        masks[("fc1", i)] = F.dropout_mask(B * S * d, p, 1 + 2 * i, step, DEV).cpu().bool().reshape(B, S, d)
        masks[("fc2", i)] = F.dropout_mask(B * S * H, p, 2 + 2 * i, step, DEV).cpu().bool().reshape(B, S, H)
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, masks=masks)
    assert abs(float(loss) - float(lo)) < 2e-5
    _grad_check(m, go, 2e-3, "own-masks")
    big = F.dropout_mask(1 << 22, 0.1, 3, 7, DEV).float()
    assert abs(float(big.mean()) - 0.9) < 2e-3
    assert torch.equal(big, F.dropout_mask(1 << 22, 0.1, 3, 7, DEV).float())
    assert not torch.equal(big, F.dropout_mask(1 << 22, 0.1, 4, 7, DEV).float())
    # a second step draws different masks
    m.zero_grad()
    loss2 = m(x.to(DEV), y.to(DEV), w)
    assert float(loss2) != float(loss)


# ----------------------------------------------------------------------------- tensor-core modes
@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("name", ["conf5", "conf9", "conf11", "conf18", "conf1", "shipped", "tiny"])
def test_forward_low_precision_within_tolerance(name, prec):
    g = load_golden(name)
    cfg, sd, m, x, y, w = build(name, prec)
    m.eval()
    with torch.no_grad():
        logits, probs, enc = m(x.to(DEV))
    tol = LOGIT_TOL[prec]
    ref = torch.from_numpy(g["logits"])
    err = float((logits.cpu() - ref).abs().max())
    assert err <= tol, (name, prec, err)
    assert enc.dtype == torch.float32 and probs[0].dtype == torch.float32
    margin = ref.abs() > 2 * tol        # classes must agree wherever the logit is not inside the tolerance band
    assert ((logits.cpu() > 0) == (ref > 0))[margin].all()


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("name", ["conf5", "conf18"])
def test_grads_low_precision(name, prec):
    cfg, sd, m, x, y, w = build(name, prec)
    m.eval()
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, dtype=torch.float64)
    tol = LOGIT_TOL[prec]
    assert abs(float(loss) - float(lo)) < tol
    _grad_check(m, {k: v.float() for k, v in go.items()}, 0.02 if prec == "tf32" else 0.08, f"{name}/{prec}",
                atol_scale=2e-4)


# ----------------------------------------------------------------------------- a9 ensemble
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_ensemble_forward_and_grads(prec):
    g = load_golden("ensemble_5_9_11")
    cfgs = [vit3d_b200.north_star_config(c) for c in (5, 9, 11)]
    members_sd = [O.init_state_dict(c, seed=42 + j) for j, c in enumerate(cfgs)]
    sd = O.ensemble_state_dict(members_sd, seed=7)
    members = [VisionTransformer(c, 128, zero_head=True, num_classes=1, precision=prec) for c in cfgs]
    ens = TransformerEnsemble(*members, in_features=1)
    ens.load_state_dict(sd)
    ens.to(DEV).eval()
    x = O.synth_volumes(3, seed=42, kind="img")
    out = ens(x.to(DEV))
    tol = 1e-4 if prec == "fp32" else 5e-3
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], atol=tol)
    assert ((out.detach().cpu() > 0.5) == (torch.from_numpy(g["out"]) > 0.5)).all()
    if prec == "fp32":
        y = O.synth_labels(3)
        loss = torch.nn.BCELoss()(out, y.unsqueeze(1).to(DEV))      # the scripts' criterion (train_ensemble_*.py:119)
        loss.backward()
        assert abs(float(loss) - float(g["loss"])) < 1e-5
        P = dict(ens.named_parameters())
        for k in ("classifier.weight", "classifier.bias", "transformers.0.head.weight",
                  "transformers.2.transformer.embeddings.patch_embeddings.weight"):
            gs = g["grad_stats/" + k]
            assert_stats_close(stats(P[k].grad), gs, 3e-3, 1e-7, k)


# ----------------------------------------------------------------------------- submodule surface (a2..a5 standalone)
def test_submodules_standalone_fp32():
    cfg = vit3d_b200.get_config(16, 96, 2, 64, 4)
    sd = O.init_state_dict(cfg, seed=11)
    torch.manual_seed(0)
    h = torch.randn(2, 65, 64)
    pre = "transformer.encoder.layer.0."
    blk = Block(cfg, True)
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    blk.precision = blk.attn.precision = blk.ffn.precision = "fp32"
    blk.to(DEV).eval()
    with torch.no_grad():
        _standalone_checks(blk, cfg, sd, h, pre)


def _standalone_checks(blk, cfg, sd, h, pre):
    out, wts = blk(h.to(DEV))
    ro, rp = O.block(sd, cfg, h, pre)
    np.testing.assert_allclose(out.cpu().numpy(), ro.numpy(), atol=2e-5)
    np.testing.assert_allclose(wts.cpu().numpy(), rp.numpy(), atol=2e-6)
    a, _ = blk.attn(h.to(DEV))
    np.testing.assert_allclose(a.cpu().numpy(), O.attention(sd, cfg, h, pre + "attn.")[0].numpy(), atol=2e-5)
    mo = blk.ffn(h.to(DEV))
    np.testing.assert_allclose(mo.cpu().numpy(), O.mlp(sd, cfg, h, pre + "ffn.").numpy(), atol=2e-5)
    emb = Embeddings(cfg, 128)
    epre = "transformer.embeddings."
    emb.load_state_dict({k[len(epre):]: v for k, v in sd.items() if k.startswith(epre)})
    emb.precision = "fp32"
    emb.to(DEV).eval()
    x = O.synth_volumes(2, seed=3)
    np.testing.assert_allclose(emb(x.to(DEV)).cpu().numpy(), O.embeddings(sd, cfg, x).numpy(), atol=2e-4, rtol=1e-5)


def test_vis_false_returns_empty_list_and_same_logits():
    cfg, sd, m, x, y, w = build("tiny", "fp32", vis=False)
    m.eval()
    with torch.no_grad():
        logits, attn, enc = m(x.to(DEV))
    assert attn == []
    np.testing.assert_allclose(logits.cpu().numpy(), load_golden("tiny")["logits"], atol=1e-4)


def test_real_volumes_predicted_classes():
    g = load_golden("real_volumes")
    u8 = torch.from_numpy(g["u8"]).float().permute(0, 4, 1, 2, 3).contiguous()
    x = u8 - u8.mean()
    cfg = vit3d_b200.north_star_config(5)
    sd = O.init_state_dict(cfg, seed=42)
    ref = torch.from_numpy(g["logits_conf5"])
    for prec in ("fp32", "bf16"):
        m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec)
        m.load_state_dict(sd)
        m.to(DEV).eval()
        with torch.no_grad():
            logits = m(x.to(DEV))[0].cpu()
        assert float((logits - ref).abs().max()) <= LOGIT_TOL[prec]
        assert ((logits > 0) == (ref > 0)).all()


def test_uint8_volume_input_equals_fp32_input():
    """N2: shipping raw uint8 volumes + the mean gives bit-identical logits to shipping (u8 - mean) fp32."""
    u8, mean = O.synth_volumes_u8(5, seed=42)
    x = O.synth_volumes(5, seed=42)
    assert torch.equal(u8.float() - mean, x)
    g = load_golden("real_volumes")
    for prec in ("fp32", "bf16"):
        cfg = vit3d_b200.north_star_config(5)
        m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec, vis=False)
        m.load_state_dict(O.init_state_dict(cfg, seed=42))
        m.to(DEV).eval()
        m.input_mean = mean
        with torch.no_grad():
            a = m(u8.to(DEV))[0]
            b = m(x.to(DEV))[0]
        assert torch.equal(a, b)
        # the real fixture: 4 volumes read through the reference's ProstateDataset, stored as uint8
        ru8 = torch.from_numpy(g["u8"]).permute(0, 4, 1, 2, 3).contiguous()
        m.input_mean = float(ru8.float().mean())
        with torch.no_grad():
            lr = m(ru8.to(DEV))[0].cpu()
        assert float((lr - torch.from_numpy(g["logits_conf5"])).abs().max()) <= LOGIT_TOL[prec]


def test_batch_sizes_and_linearity_property():
    """Size-independent property at a larger batch: volumes are independent, so a batch's logits equal
    the per-volume logits (checks the flat token-matrix tiling at ragged row counts)."""
    cfg = vit3d_b200.north_star_config(5)
    sd = O.init_state_dict(cfg, seed=42)
    for prec, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec, vis=False)
        m.load_state_dict(sd)
        m.to(DEV).eval()
        x = O.synth_volumes(37, seed=8).to(DEV)
        with torch.no_grad():
            full = m(x)[0]
            parts = torch.cat([m(x[i:i + 5])[0] for i in range(0, 37, 5)])
            one = m(x[36:37])[0]
        assert float((full - parts).abs().max()) <= tol
        assert float((full[36:] - one).abs().max()) <= tol


def test_full_size_batch_properties():
    """BASELINE.json's bench configuration itself (conf 5, batch 1024, bf16, vis=True): the oracle cannot run 1024
    volumes in seconds, so the full-size run is checked through size-independent properties - volumes are
    independent (the batch's logits equal those of its 64-volume chunks), every softmax row of every layer sums
    to 1, the encoder output is LayerNorm-ed (finite, unit scale) - and a sample of volumes against the oracle."""
    cfg = vit3d_b200.north_star_config(5)
    sd = O.init_state_dict(cfg, seed=42)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16", vis=True)
    m.load_state_dict(sd)
    m.to(DEV).eval()
    B = 1024
    x = O.synth_volumes(B, seed=42)
    xd = x.to(DEV)
    with torch.no_grad():
        logits, attn, enc = m(xd)
        assert logits.shape == (B, 1) and enc.shape == (B, 65, 256) and len(attn) == 6
        assert attn[0].shape == (B, 8, 65, 65)
        for a in (attn[0], attn[-1]):
            rs = a.sum(-1)
            assert float((rs - 1).abs().max()) <= 1e-3 and float(a.min()) >= 0.0
        assert torch.isfinite(enc).all() and torch.isfinite(logits).all()
        logits, enc0 = logits.clone(), enc[:, 0].clone()
        del attn, enc
        parts = torch.cat([m(xd[i:i + 64])[0] for i in range(0, B, 64)])
    assert float((logits - parts).abs().max()) <= 1e-5          # same rows, same arithmetic, other tiles
    pick = [0, 1, 63, 64, 511, 777, 1022, 1023]
    ref_logits, _, ref_enc = O.vit_forward(sd, cfg, x[pick])
    assert float((logits[pick].cpu() - ref_logits).abs().max()) <= LOGIT_TOL["bf16"]
    clear = ref_logits.abs() > LOGIT_TOL["bf16"]                       # volume 0 sits at 0.002: inside the tolerance band
    assert torch.equal((logits[pick].cpu() > 0)[clear], (ref_logits > 0)[clear])   # identical predicted classes
    assert float((enc0[pick].cpu() - ref_enc[:, 0]).abs().max()) <= 0.05 * float(ref_enc.abs().max())


def test_full_size_training_step_linearity():
    """BASELINE.json's training configuration (conf 18, batch 256, bf16): BCE is a mean over the batch, so with a
    fixed pos_weight the loss and every parameter gradient of the full batch equal the average over its two halves
    (eval mode: no dropout noise).  Checks the backward kernels at the bench's tile counts."""
    cfg = vit3d_b200.north_star_config(18)
    sd = O.init_state_dict(cfg, seed=42)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16", vis=False)
    m.load_state_dict(sd)
    m.to(DEV).eval()
    B = 256
    x = O.synth_volumes(B, seed=7).to(DEV)
    y = O.synth_labels(B).to(DEV)
    w = 1.7

    def run(xs, ys):
        m.zero_grad(set_to_none=True)
        loss = m(xs, ys, w)
        loss.backward()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    lf, gf = run(x, y)
    la, ga = run(x[:128], y[:128])
    lb, gb = run(x[128:], y[128:])
    assert np.isfinite(lf) and abs(lf - 0.5 * (la + lb)) <= 2e-3 * max(1.0, abs(lf))
    worst = 0.0
    for k in gf:
        ref = 0.5 * (ga[k] + gb[k])
        scale = float(ref.abs().max()) + 1e-12
        worst = max(worst, float((gf[k] - ref).abs().max()) / scale)
        assert torch.isfinite(gf[k]).all(), k
    # bf16 operands: the halves round their activations exactly like the full batch (rows are independent), what
    # differs is the summation order of the weight-gradient reductions (split-K atomics)
    assert worst <= 2e-2, worst


# ----------------------------------------------------------------------------- N1 optimizers
def test_fused_sgd_and_adam_match_torch():
    from vit3d_b200._lib import call, ptr, stream
    torch.manual_seed(1)
    n = 10007
    p0, g1, g2 = torch.randn(n), torch.randn(n), torch.randn(n)
    # SGD(momentum .9, wd 1e-2, lr 1e-4): train_baseline_cv.py:111-114
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.9, weight_decay=1e-2)
    p = p0.clone().to(DEV)
    mom = torch.zeros(n, device=DEV)
    for i, g in enumerate((g1, g2)):
        pr.grad = g.clone()
        opt.step()
        call("vit3d_sgd_step", ptr(p), ptr(g.to(DEV)), ptr(mom), n, 1e-2, 0.9, 1e-2, int(i == 0), 1.0, None, stream())
    np.testing.assert_allclose(p.cpu().numpy(), pr.detach().numpy(), atol=1e-6)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pr], lr=1e-3)
    p = p0.clone().to(DEV)
    m_, v_ = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for i, g in enumerate((g1, g2)):
        pr.grad = g.clone()
        opt.step()
        call("vit3d_adam_step", ptr(p), ptr(g.to(DEV)), ptr(m_), ptr(v_), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, i + 1, 1.0,
             None, None, stream())
    np.testing.assert_allclose(p.cpu().numpy(), pr.detach().numpy(), atol=2e-6)


def test_state_dict_roundtrip_on_device():
    cfg, sd, m, x, y, w = build("tiny", "fp32")
    out = {k: v.cpu() for k, v in m.state_dict().items()}
    assert sorted(out) == sorted(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k])

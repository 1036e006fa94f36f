"""Unit parity of the fused bf16 attention kernel (mma.sync + bulk-async staging) against an fp64
softmax attention over the same bf16-rounded qkv."""
import math

import numpy as np
import pytest
import torch

from vit3d_b200._lib import PREC, call, lib, ptr, stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ref_attention(qkv, heads):
    B, S, A3 = qkv.shape
    A = A3 // 3
    D = A // heads
    q, k, v = (qkv[..., i * A:(i + 1) * A].double().view(B, S, heads, D).permute(0, 2, 1, 3) for i in range(3))
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(D), dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, S, A), p


@pytest.mark.parametrize("heads", [4, 8, 16])
@pytest.mark.parametrize("B", [1, 3, 149, 700])
@pytest.mark.parametrize("vis", [True, False])
@pytest.mark.parametrize("threads", [0, 640, 512, "unit"])
def test_attention_bf16(heads, B, vis, threads):
    """threads: CTA size of the one-volume-per-CTA kernel, or "unit" = the (volume, 4-head) unit kernel."""
    unit = threads == "unit"
    threads = 0 if unit else threads
    lib().vit3d_set_tuning(14, 2 if unit else 0)
    torch.manual_seed(B * 31 + heads)
    S, A = 65, 256
    qkv = (torch.randn(B, S, 3 * A, device=DEV) * 1.5).to(torch.bfloat16)
    ctx = torch.full((B, S, A), float("nan"), device=DEV, dtype=torch.bfloat16)
    probs = torch.full((B, heads, S, S), float("nan"), device=DEV) if vis else None
    lib().vit3d_set_tuning(1, threads)
    try:
        call("vit3d_attn_fwd", ptr(qkv), ptr(ctx), ptr(probs), B, S, heads, A // heads, PREC["bf16"], stream())
        torch.cuda.synchronize()
    finally:
        lib().vit3d_set_tuning(1, 0)
        lib().vit3d_set_tuning(14, 0)
    rc, rp = ref_attention(qkv, heads)
    err = float((ctx.double() - rc).abs().max())
    assert np.isfinite(err) and err < 0.03, err                 # bf16 P and bf16 output rounding
    if vis:
        perr = float((probs.double() - rp).abs().max())
        assert np.isfinite(perr) and perr < 2e-5, perr         # softmax itself is fp32
        assert float((probs.sum(-1) - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("heads", [4, 8, 16])
@pytest.mark.parametrize("B", [1, 3, 149, 700])
@pytest.mark.parametrize("vis", [True, False])
def test_attention_tf32_mode(heads, B, vis):
    """TF32-mode attention (fp32 q | k | v, mma.sync tf32 in (volume, 4-head) units) against fp64 softmax attention, and
    against the generic fp32 SIMT kernel it replaces: fp32-grade probabilities (3xTF32 scores), a context that differs
    by the tf32 rounding of P and V."""
    torch.manual_seed(B * 29 + heads)
    S, A = 65, 256
    qkv = torch.randn(B, S, 3 * A, device=DEV) * 1.5
    out = {}
    for kern in (1, 0):
        ctx = torch.full((B, S, A), float("nan"), device=DEV)
        probs = torch.full((B, heads, S, S), float("nan"), device=DEV) if vis else None
        lib().vit3d_set_tuning(9, kern)
        try:
            call("vit3d_attn_fwd", ptr(qkv), ptr(ctx), ptr(probs), B, S, heads, A // heads, PREC["tf32"], stream())
            torch.cuda.synchronize()
        finally:
            lib().vit3d_set_tuning(9, 1)
        out[kern] = (ctx, probs)
    rc, rp = ref_attention(qkv, heads)
    ctx, probs = out[1]
    err = float((ctx.double() - rc).abs().max())
    scale = float(rc.abs().max())
    assert np.isfinite(err) and err < 1.5e-3 * scale, (err, scale)   # P and V rounded to tf32 (2^-11 relative each)
    rel = float((ctx.double() - rc).norm() / rc.norm())
    assert rel < 5e-4, rel
    assert float((ctx - out[0][0]).abs().max()) < 1.5e-3 * scale
    if vis:
        perr = float((probs.double() - rp).abs().max())
        assert np.isfinite(perr) and perr < 2e-5, perr         # scores are a 3xTF32 product: fp32-grade probabilities
        assert float((probs.sum(-1) - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("kern", [1, 0])
@pytest.mark.parametrize("heads", [4, 8, 16])
@pytest.mark.parametrize("B", [1, 3, 149, 700])
def test_attention_backward_bf16(heads, B, kern):
    """mma.sync attention backward (probabilities recomputed) vs fp64 autograd on the same bf16 inputs: the unit kernel
    (kern=1: (volume, 4-head) units, TMA boxes) and the one-volume-per-CTA kernel (kern=0), with the q | k | v bias
    gradients (column sums of dqkv) accumulated onto existing values."""
    from vit3d_b200._lib import lib
    torch.manual_seed(B * 17 + heads)
    S, A = 65, 256
    qkv = (torch.randn(B, S, 3 * A, device=DEV) * 1.2).to(torch.bfloat16)
    dctx = torch.randn(B, S, A, device=DEV).to(torch.bfloat16)
    dqkv = torch.full((B, S, 3 * A), float("nan"), device=DEV, dtype=torch.bfloat16)
    db = [torch.full((A,), 0.5, device=DEV) for _ in range(3)]
    lib().vit3d_set_tuning(7, kern)
    try:
        call("vit3d_attn_bwd_bias", ptr(dctx), ptr(qkv), ptr(dqkv), ptr(db[0]), ptr(db[1]), ptr(db[2]), B, S, heads, A // heads, stream())
        torch.cuda.synchronize()
        plain = torch.full((B, S, 3 * A), float("nan"), device=DEV, dtype=torch.bfloat16)
        call("vit3d_attn_bwd", ptr(dctx), ptr(qkv), ptr(plain), B, S, heads, A // heads, PREC["bf16"], stream())
        torch.cuda.synchronize()
    finally:
        lib().vit3d_set_tuning(7, 1)
    assert torch.equal(plain, dqkv)
    q64 = qkv.double().requires_grad_(True)
    ctx, _ = ref_attention(q64, heads)
    ctx.backward(dctx.double())
    ref = q64.grad
    err = float((dqkv.double() - ref).abs().max())
    scale = float(ref.abs().max())
    assert np.isfinite(err) and err <= 0.03 * scale, (err, scale)
    rel = float((dqkv.double() - ref).norm() / ref.norm())
    assert rel < 0.01, rel
    got_b = torch.cat(db).double() - 0.5
    ref_b = dqkv.double().sum((0, 1))
    assert float((got_b - ref_b).abs().max()) <= 2e-3 * float(ref_b.abs().max()) + 1e-3 * B ** 0.5


@pytest.mark.parametrize("heads", [4, 8, 16])
def test_padded_probability_rows_equal_packed_rows(heads):
    """vit3d_attn_fwd_padded (rows of 72 floats, sector-aligned stores) writes the same probabilities and context as the
    packed layout of the reference tensor; the padding floats are zeros."""
    import vit3d_b200  # noqa: F401
    from vit3d_b200._lib import PREC, call, ptr, stream
    B, S, A = 150, 65, 256
    D = A // heads
    torch.manual_seed(heads)
    qkv = (torch.randn(B, S, 3 * A, device=DEV) * 0.7).to(torch.bfloat16)
    ctx0 = torch.empty(B, S, A, device=DEV, dtype=torch.bfloat16)
    ctx1 = torch.empty_like(ctx0)
    packed = torch.empty(B, heads, S, S, device=DEV)
    padded = torch.full((B, heads, S, 72), -7.0, device=DEV)
    call("vit3d_attn_fwd", ptr(qkv), ptr(ctx0), ptr(packed), B, S, heads, D, PREC["bf16"], stream())
    call("vit3d_attn_fwd_padded", ptr(qkv), ptr(ctx1), ptr(padded), 72, B, S, heads, D, stream())
    assert torch.equal(padded[..., :S], packed)
    assert torch.equal(ctx0, ctx1)
    assert bool((padded[..., S:] == 0.0).all())
    assert abs(float(padded[..., :S].sum(-1).mean()) - 1.0) < 1e-5

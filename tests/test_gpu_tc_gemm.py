"""Unit parity of the tcgen05 GEMM (vit3d_linear_fwd in TF32 / BF16 modes) against an fp64 product of
the same (rounded) operands, over the shapes the model uses plus ragged edges."""
import numpy as np
import pytest
import torch

import vit3d_b200
from vit3d_b200._lib import PREC, call, lib, ptr, stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    # M, N, K
    (130, 64, 64), (260, 768, 256), (1300, 2048, 256), (1300, 256, 2048), (129, 256, 256), (65, 3072, 256),
    (1000, 72, 96), (4160, 256, 1280), (16640, 2048, 256), (333, 200, 264),
]


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / 2 ** 0.5))


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["plain", "bias_gelu_pre", "bias_residual"])
def test_tc_linear_matches_fp64(prec, shape, mode):
    M, N, K = shape
    torch.manual_seed(M * 7 + N * 3 + K)
    lp = prec == "bf16"
    adt = torch.bfloat16 if lp else torch.float32
    x = torch.randn(M, K, device=DEV).to(adt)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5)
    wl = w.to(torch.bfloat16) if lp else None
    b = torch.randn(N, device=DEV) if mode != "plain" else None
    res = torch.randn(M, N, device=DEV) if mode == "bias_residual" else None
    act = 1 if mode == "bias_gelu_pre" else 0
    yf = (res is not None) or not lp
    y = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32 if yf else torch.bfloat16)
    pre = torch.full_like(y, float("nan")) if act else None
    assert lib().vit3d_tc_supported(PREC[prec], M, N, K) == 1
    call("vit3d_linear_fwd", ptr(x), K, int(not lp), ptr(w), ptr(wl), ptr(b), ptr(res), ptr(y), int(yf), ptr(pre), act,
         M, N, K, PREC[prec], stream())
    torch.cuda.synchronize()
    xw = x.double() @ (wl.double() if lp else w.double()).t()
    if b is not None:
        xw = xw + b.double()
    ref_pre = xw
    ref = gelu(xw) if act else xw
    if res is not None:
        ref = ref + res.double()
    # bf16 operands are exact in the reference; what is left is fp32 accumulation (+ tf32 operand rounding, + bf16 output rounding)
    tol = (2e-5 if lp else 2e-3) * (K ** 0.5) * 0.5 + (0.0 if yf else 1.0 / 128) * float(ref.abs().max())
    err = float((y.double() - ref).abs().max())
    assert np.isfinite(err) and err <= tol, (prec, shape, mode, err, tol)
    if pre is not None:
        errp = float((pre.double() - ref_pre).abs().max())
        assert errp <= tol, (errp, tol)


@pytest.mark.parametrize("variant", ["cluster4", "cluster2", "tall", "plain"])
@pytest.mark.parametrize("B", [1, 2, 3, 37, 597, 600, 1024])
@pytest.mark.parametrize("H", [256, 64])
def test_tc_patch_embedding_matches_oracle(B, H, variant):
    """TMA-im2col TF32 patch embedding (+bias, +position rows, cls rows) vs the fp64 oracle.  Batches of >= 592 volumes
    can run on clusters of 4 / 2 CTAs sharing the filter bank by TMA multicast (taken when the tile count lets every CTA
    of a cluster walk equally many tiles: 600 and 1024 volumes, not 597) or on 256-row tiles (597 volumes leave the
    last tile's second half ragged)."""
    if B == 1024 and (H != 256 or variant in ("tall", "plain")):
        pytest.skip("large case only for the cluster variants")
    lib().vit3d_set_tuning(11, 1 if variant == "tall" else 0)
    lib().vit3d_set_tuning(12, {"cluster4": 4, "cluster2": 2}.get(variant, 0))
    from oracle import vit3d_oracle as O
    from vit3d_b200.models.modeling import Embeddings
    cfg = vit3d_b200.get_config(16, 128, 1, H, 4)
    sd = O.init_state_dict(cfg, seed=5)
    pre = "transformer.embeddings."
    emb = Embeddings(cfg, 128)
    emb.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    emb.precision = "bf16"
    emb.to(DEV).eval()
    x = O.synth_volumes(B, seed=B)
    try:
        with torch.no_grad():
            got = emb(x.to(DEV)).cpu().double()
    finally:
        lib().vit3d_set_tuning(11, 0)
        lib().vit3d_set_tuning(12, 0)
    ref = O.embeddings({k: v.double() for k, v in sd.items()}, cfg, x.double())
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    assert got.shape == ref.shape and np.isfinite(err)
    assert err <= 2e-3 * scale, (err, scale)       # tf32 operands (10-bit mantissa), K = 1280
    # cls rows are exact fp32 adds
    cls_ref = (sd[pre + "cls_token"][0, 0] + sd[pre + "position_embeddings"][0, 0]).double()
    assert float((got[:, 0] - cls_ref).abs().max()) < 1e-6


@pytest.mark.parametrize("shape", [(1300, 768, 256), (16640, 2048, 256), (16640, 256, 2048), (260, 256, 256),
                                   (4161, 3072, 256), (130, 64, 64)])
def test_tc_linear_backward_bf16(shape):
    """tcgen05 data gradient (transposed bf16 weight shadow) and weight gradient (both operands MN-major,
    split over the token rows, fp32 atomics) vs fp64 products of the same bf16 operands."""
    M, N, K = shape
    torch.manual_seed(M + N + K)
    dy = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    wt = torch.empty(K, N, device=DEV, dtype=torch.bfloat16)
    call("vit3d_transpose_f32_to_bf16", ptr(w), ptr(wt), N, K, stream())
    assert torch.equal(wt, w.t().contiguous().to(torch.bfloat16))
    for dx_f32 in (0, 1):
        dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.float32 if dx_f32 else torch.bfloat16)
        dw = torch.zeros(N, K, device=DEV)
        db = torch.zeros(N, device=DEV)
        call("vit3d_linear_bwd", ptr(dy), 0, ptr(x), K, 0, ptr(w), ptr(wt), ptr(dx), K, dx_f32, ptr(dw), ptr(db), M, N, K,
             PREC["bf16"], stream())
        torch.cuda.synchronize()
        ref_dx = dy.double() @ wt.double().t()
        ref_dw = dy.double().t() @ x.double()
        ref_db = dy.double().sum(0)
        e = float((dx.double() - ref_dx).abs().max())
        assert np.isfinite(e) and e <= 1e-4 * N ** 0.5 + (0.0 if dx_f32 else 1.0 / 128) * float(ref_dx.abs().max()), e
        e = float((dw.double() - ref_dw).abs().max())
        assert np.isfinite(e) and e <= 2e-5 * M ** 0.5 * 4, (e, float(ref_dw.abs().max()))
        assert float((db.double() - ref_db).abs().max()) <= 1e-3 * M ** 0.5


@pytest.mark.parametrize("M", [65, 130, 260, 4160, 19000])
@pytest.mark.parametrize("d", [2048, 3072, 256])
def test_fused_mlp_matches_fp64(M, d):
    """Single-kernel fc1 -> GELU -> fc2 (+bias, +residual) vs an fp64 evaluation on the same bf16 operands
    (the bf16 rounding of the GELU output is reproduced in the reference)."""
    torch.manual_seed(M + d)
    H = 256
    xn = torch.randn(M, H, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(d, H, device=DEV) / H ** 0.5)
    w2 = (torch.randn(H, d, device=DEV) / d ** 0.5)
    b1 = torch.randn(d, device=DEV) * 0.1
    b2 = torch.randn(H, device=DEV) * 0.1
    res = torch.randn(M, H, device=DEV)
    w1l, w2l = w1.to(torch.bfloat16), w2.to(torch.float16)
    out = torch.full((M, H), float("nan"), device=DEV)
    assert lib().vit3d_mlp_supported(M, H, d) == 1
    call("vit3d_mlp_fwd", ptr(xn), ptr(w1l), ptr(b1), ptr(w2l), ptr(b2), ptr(res), ptr(out), M, H, d, stream())
    torch.cuda.synchronize()
    h = xn.double() @ w1l.double().t() + b1.double()
    a = gelu(h).to(torch.float16).double()             # the kernel keeps GELU(h) in 16 bits before fc2
    ref = a @ w2l.double().t() + b2.double() + res.double()
    err = float((out.double() - ref).abs().max())
    # a bf16 ulp flip of one GELU output (value ~ up to 4) moves one product by <= 2^-7 * |w2| ~ 1e-3; many flip
    assert np.isfinite(err) and err <= 0.02, err
    rel = float((out.double() - ref).norm() / ref.norm())
    assert rel < 2e-3, rel


@pytest.mark.parametrize("shape", [(65, 256, 256), (130, 256, 256), (4161, 256, 256), (1300, 256, 2048), (16640, 256, 3072),
                                   (19000, 256, 256)])
@pytest.mark.parametrize("alias", [False, True])
def test_linear_residual_layernorm_fused(shape, alias):
    """vit3d_linear_ln_fwd: y = x W^T + b + residual (fp32) and LN(y) (bf16) from one GEMM epilogue, against
    an fp64 evaluation on the same bf16 operands; mean / rstd outputs; y may alias the residual."""
    M, N, K = shape
    torch.manual_seed(M + K)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    wl = w.to(torch.bfloat16)
    b = torch.randn(N, device=DEV) * 0.1
    res = torch.randn(M, N, device=DEV) * 2.0 + 0.5
    gamma = 1.0 + 0.1 * torch.randn(N, device=DEV)
    beta = 0.1 * torch.randn(N, device=DEV)
    eps = 1e-6
    res_in = res.clone()
    y = res_in if alias else torch.full((M, N), float("nan"), device=DEV)
    yn = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    mean = torch.full((M,), float("nan"), device=DEV)
    rstd = torch.full((M,), float("nan"), device=DEV)
    assert lib().vit3d_linear_ln_supported(M, N, K) == 1
    call("vit3d_linear_ln_fwd", ptr(x), ptr(wl), ptr(b), ptr(res_in), ptr(y), ptr(gamma), ptr(beta), eps, ptr(yn),
         ptr(mean), ptr(rstd), M, N, K, stream())
    torch.cuda.synchronize()
    ref = x.double() @ wl.double().t() + b.double() + res.double()
    mu = ref.mean(1, keepdim=True)
    var = ref.var(1, unbiased=False, keepdim=True)
    ref_n = (ref - mu) / torch.sqrt(var + eps) * gamma.double() + beta.double()
    e = float((y.double() - ref).abs().max())
    assert np.isfinite(e) and e <= 2e-5 * K ** 0.5, e
    assert float((mean.double() - mu[:, 0]).abs().max()) <= 1e-5
    assert float((rstd.double() * torch.sqrt(var[:, 0] + eps) - 1).abs().max()) <= 1e-4
    en = float((yn.double() - ref_n).abs().max())
    assert np.isfinite(en) and en <= float(ref_n.abs().max()) / 128, en      # bf16 output rounding
    # unsupported shapes are refused, not mis-computed
    assert lib().vit3d_linear_ln_supported(M, 512, K) == 0


@pytest.mark.parametrize("conf", [5, 18])
def test_inference_ln_chain_matches_unfused(conf, monkeypatch):
    """The inference fast path (LayerNorm fused into the out-proj / fc2 epilogues, cached packed q|k|v weight)
    against the operator-by-operator composition the training path uses, same weights, same input."""
    from oracle import vit3d_oracle as O
    from vit3d_b200.models.modeling import VisionTransformer
    cfg = vit3d_b200.north_star_config(conf)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16")
    m.load_state_dict(O.init_state_dict(cfg, seed=7))
    m.to(DEV).eval()
    x = O.synth_volumes(6, seed=3).to(DEV)
    with torch.no_grad():
        lf, af, ef = m(x)
    with torch.enable_grad():          # grad mode on -> the unfused composition (nothing requires grad on x)
        lu, au, eu = m(x)
    assert float((lf - lu).abs().max()) <= 1.5e-2      # two bf16-mode evaluations, each within ~6e-3 of the fp32 truth
    assert float((ef - eu).abs().max()) <= 0.05 * float(eu.abs().max())
    assert len(af) == len(au) and float((af[-1] - au[-1]).abs().max()) <= 2e-2


@pytest.mark.parametrize("M", [65, 130, 260, 4160, 19000])
@pytest.mark.parametrize("d", [2048, 3072, 256])
@pytest.mark.parametrize("pair", [0, 1, 2, 3])
def test_fused_mlp_layernorm_matches_fp64(M, d, pair):
    """vit3d_mlp_ln_fwd (chunked fused MLP, single CTA and CTA pair): y = x + fc2(GELU(fc1(xn))) in fp32 and
    LayerNorm(y) in bf16 vs an fp64 evaluation on the same operands; y written in place over the residual."""
    torch.manual_seed(M + d + pair)
    H = 256
    xn = torch.randn(M, H, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(d, H, device=DEV) / H ** 0.5)
    w2 = (torch.randn(H, d, device=DEV) / d ** 0.5)
    b1 = torch.randn(d, device=DEV) * 0.1
    b2 = torch.randn(H, device=DEV) * 0.1
    res = torch.randn(M, H, device=DEV)
    gamma = 1.0 + 0.1 * torch.randn(H, device=DEV)
    beta = 0.1 * torch.randn(H, device=DEV)
    w1l, w2l = w1.to(torch.bfloat16), w2.to(torch.float16)
    out = res.clone()                      # in place
    yn = torch.full((M, H), float("nan"), device=DEV, dtype=torch.bfloat16)
    assert lib().vit3d_mlp_ln_supported(M, H, d) == 1
    lib().vit3d_set_tuning(5, pair)
    try:
        call("vit3d_mlp_ln_fwd", ptr(xn), ptr(w1l), ptr(b1), ptr(w2l), ptr(b2), ptr(out), ptr(out), ptr(gamma), ptr(beta),
             1e-6, ptr(yn), M, H, d, stream())
        torch.cuda.synchronize()
    finally:
        lib().vit3d_set_tuning(5, 0)
    h = xn.double() @ w1l.double().t() + b1.double()
    a = gelu(h).to(torch.float16).double()
    ref = a @ w2l.double().t() + b2.double() + res.double()
    err = float((out.double() - ref).abs().max())
    assert np.isfinite(err) and err <= 0.02, err
    assert float((out.double() - ref).norm() / ref.norm()) < 2e-3
    mu = ref.mean(1, keepdim=True)
    var = ref.var(1, unbiased=False, keepdim=True)
    ref_n = (ref - mu) / torch.sqrt(var + 1e-6) * gamma.double() + beta.double()
    en = float((yn.double() - ref_n).abs().max())
    assert np.isfinite(en) and en <= 0.02 + float(ref_n.abs().max()) / 128, en


@pytest.mark.parametrize("M", [65, 130, 4160, 19000, 66560])
@pytest.mark.parametrize("d", [2048, 3072])
@pytest.mark.parametrize("pair", [0, 1, 2, 3])
def test_fused_mlp_final_layernorm_fp32(M, d, pair):
    """vit3d_mlp_lnf_fwd (last Block: encoder_norm folded into the MLP kernel, fp32 output only) vs an fp64
    evaluation on the same operands, and vs LayerNorm of the y that vit3d_mlp_ln_fwd writes; the residual
    buffer must be left untouched (the kernel has no y output)."""
    torch.manual_seed(M + d + pair + 1)
    H = 256
    xn = torch.randn(M, H, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(d, H, device=DEV) / H ** 0.5)
    w2 = (torch.randn(H, d, device=DEV) / d ** 0.5)
    b1 = torch.randn(d, device=DEV) * 0.1
    b2 = torch.randn(H, device=DEV) * 0.1
    res = torch.randn(M, H, device=DEV)
    gamma = 1.0 + 0.1 * torch.randn(H, device=DEV)
    beta = 0.1 * torch.randn(H, device=DEV)
    w1l, w2l = w1.to(torch.bfloat16), w2.to(torch.float16)
    res_in = res.clone()
    enc = torch.full((M, H), float("nan"), device=DEV)
    y = torch.empty(M, H, device=DEV)
    yn = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    lib().vit3d_set_tuning(5, pair)
    try:
        call("vit3d_mlp_lnf_fwd", ptr(xn), ptr(w1l), ptr(b1), ptr(w2l), ptr(b2), ptr(res_in), ptr(gamma), ptr(beta), 1e-6,
             ptr(enc), M, H, d, stream())
        call("vit3d_mlp_ln_fwd", ptr(xn), ptr(w1l), ptr(b1), ptr(w2l), ptr(b2), ptr(res), ptr(y), ptr(gamma), ptr(beta),
             1e-6, ptr(yn), M, H, d, stream())
        torch.cuda.synchronize()
    finally:
        lib().vit3d_set_tuning(5, 0)
    assert torch.equal(res_in, res)
    # same accumulators, same statistics: LayerNorm of the other kernel's fp32 y agrees to fp32 rounding
    ln_y = torch.nn.functional.layer_norm(y.double(), (H,), gamma.double(), beta.double(), 1e-6)
    e1 = float((enc.double() - ln_y).abs().max())
    assert np.isfinite(e1) and e1 <= 2e-5 * max(1.0, float(ln_y.abs().max())), e1
    if M <= 19000:
        h = xn.double() @ w1l.double().t() + b1.double()
        a = gelu(h).to(torch.float16).double()
        ref = a @ w2l.double().t() + b2.double() + res.double()
        ref_n = torch.nn.functional.layer_norm(ref, (H,), gamma.double(), beta.double(), 1e-6)
        en = float((enc.double() - ref_n).abs().max())
        assert en <= 0.02, en

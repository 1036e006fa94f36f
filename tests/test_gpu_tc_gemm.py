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

"""Fused BF16 training step (fused_train.py) against the oracle (reference `loss = model(x, y, w); loss.backward()`,
train_baseline_cv.py:171-176): eval-mode gradients, training mode with the reference's own dropout masks injected,
training mode with the library's Philox masks, the per-operator path as a second opinion, and the CUDA-graph'd step.

Tolerance: bf16 operands -> parameter gradients within 8 % relative L2 norm (measured 1-4 %), loss within 2e-2."""
import numpy as np
import pytest
import torch

import vit3d_b200
from oracle import vit3d_oracle as O
from tests.helpers import case_setup, load_golden, unpack_masks
from vit3d_b200 import _lib, functional as F, fused_train
from vit3d_b200.graphs import GraphedInference, GraphedTrainStep
from vit3d_b200.optim import FusedSGD
from vit3d_b200.models.modeling import VisionTransformer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_RTOL = 0.08


def build(name, dropout=None):
    cfg, sd, x, y, w = case_setup(name)
    if dropout is not None:
        cfg.transformer["dropout_rate"] = dropout
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16")
    m.load_state_dict(sd)
    return cfg, sd, m.to(DEV), x, y, w


def grad_check(m, go, rtol, what, atol_scale=2e-4):
    gscale = max(float(v.norm()) for v in go.values())
    worst = 0.0
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        a, b = p.grad.detach().float().cpu(), go[k].float()
        err, ref = float((a - b).norm()), float(b.norm())
        assert err <= rtol * ref + atol_scale * gscale, (what, k, err, ref)
        worst = max(worst, err / (ref + atol_scale * gscale))
    return worst


def site_masks(cfg, masks):
    sites = {0: masks["emb"].reshape(-1)}
    for i in range(cfg.transformer["num_layers"]):
        sites[1 + 2 * i] = masks[("fc1", i)].reshape(-1)
        sites[2 + 2 * i] = masks[("fc2", i)].reshape(-1)
    return sites


@pytest.mark.parametrize("name", ["conf5", "conf18", "conf1"])
def test_fused_step_is_taken_and_matches_oracle_eval_mode(name):
    cfg, sd, m, x, y, w = build(name)
    m.eval()
    assert fused_train.supported(m, x.to(DEV))
    n0 = _lib.lib().vit3d_launch_count()
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    launches = _lib.lib().vit3d_launch_count() - n0
    L = cfg.transformer["num_layers"]
    assert launches <= 17 * L + 20, launches            # 5 forward + 12 backward launches per Block
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, dtype=torch.float64)
    assert abs(float(loss) - float(lo)) < 2e-2
    grad_check(m, go, GRAD_RTOL, name)


def test_fused_step_with_reference_dropout_masks():
    """Training mode, the masks the unmodified reference drew (tests/golden/conf5.npz) injected as keep bits."""
    g = load_golden("conf5")
    cfg, sd, m, x, y, w = build("conf5")
    masks = unpack_masks(g)
    m.train()
    with F.mask_injection(site_masks(cfg, masks)):
        loss = m(x.to(DEV), y.to(DEV), w)
        loss.backward()
    assert abs(float(loss) - float(g["loss_train"])) < 2e-2
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, masks=masks, dtype=torch.float64)
    grad_check(m, go, GRAD_RTOL, "conf5/ref-masks")


def test_fused_step_own_masks_equal_dropout_mask_export():
    """The keep bits vit3d_dropout_bits draws are the masks vit3d_dropout_mask exports for the same (seed, site,
    step): the oracle with those masks reproduces the fused training step; a second step draws new masks."""
    cfg, sd, m, x, y, w = build("conf5")
    m.train()
    torch.manual_seed(1234)
    step = F._STATE["step"] + 1
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    B, S, H, d = x.shape[0], 65, cfg.hidden_size, cfg.transformer["mlp_dim"]
    p = cfg.transformer["dropout_rate"]
    masks = {"emb": F.dropout_mask(B * S * H, p, 0, step, DEV).cpu().bool().reshape(B, S, H)}
    for i in range(cfg.transformer["num_layers"]):
        masks[("fc1", i)] = F.dropout_mask(B * S * d, p, 1 + 2 * i, step, DEV).cpu().bool().reshape(B, S, d)
        masks[("fc2", i)] = F.dropout_mask(B * S * H, p, 2 + 2 * i, step, DEV).cpu().bool().reshape(B, S, H)
    lo, go, _ = O.vit_loss_and_grads(sd, cfg, x, y, w, masks=masks, dtype=torch.float64)
    assert abs(float(loss) - float(lo)) < 2e-2
    grad_check(m, go, GRAD_RTOL, "conf5/own-masks")
    m.zero_grad()
    loss2 = m(x.to(DEV), y.to(DEV), w)
    assert float(loss2) != float(loss)


def test_dropout_bits_layout_and_rate():
    """Bit e of word w = element 32 w + e; segments are laid end to end; the keep rate is 1 - p."""
    n = [1 << 20, 1 << 18]
    import ctypes as C
    bits = torch.empty(sum(n) // 8, device=DEV, dtype=torch.uint8)
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    _lib.call("vit3d_dropout_bits", bits.data_ptr(), 2, (C.c_uint * 2)(3, 8), (C.c_longlong * 2)(*n), 0.1, seed, 5, None,
              torch.cuda.current_stream().cuda_stream)
    unpacked = ((bits.view(-1, 1).to(torch.int32) >> torch.arange(8, device=DEV, dtype=torch.int32)) & 1).reshape(-1)
    assert torch.equal(unpacked[:n[0]].to(torch.uint8), F.dropout_mask(n[0], 0.1, 3, 5, DEV))
    assert torch.equal(unpacked[n[0]:].to(torch.uint8), F.dropout_mask(n[1], 0.1, 8, 5, DEV))
    assert abs(float(unpacked.float().mean()) - 0.9) < 2e-3


def test_fused_step_agrees_with_per_operator_path():
    """Second opinion: the same bf16 arithmetic composed from the per-operator autograd Functions."""
    cfg, sd, m, x, y, w = build("conf18")
    m.eval()
    loss = m(x.to(DEV), y.to(DEV), w)
    loss.backward()
    fused = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    F._STATE["fused_train"] = False
    try:
        loss_op = m(x.to(DEV), y.to(DEV), w)
        loss_op.backward()
    finally:
        F._STATE["fused_train"] = True
    assert abs(float(loss) - float(loss_op)) < 5e-3
    gmax = max(float(v.norm()) for v in fused.values())
    for k, p in m.named_parameters():
        err = float((p.grad - fused[k]).norm())
        assert err <= 0.05 * float(fused[k].norm()) + 2e-4 * gmax, (k, err)


def test_fused_step_into_flat_arena_accumulates_and_ragged_batches():
    """Gradients accumulate (+=) straight into the optimizer's flat arena; batch sizes that do not fill 128-row
    tiles (1, 3, 5 volumes) agree with the oracle."""
    cfg, sd, m, x, y, w = build("conf5")
    m.eval()
    opt = FusedSGD(m.parameters(), lr=0.0)
    for B in (1, 3, 5):
        xb = O.synth_volumes(B, seed=5)
        yb = O.synth_labels(max(B, 2))[:B]
        opt.zero_grad()
        for _ in range(2):                       # two backward passes: the arena holds their sum
            loss = m(xb.to(DEV), yb.to(DEV), 1.3)
            loss.backward()
        assert all(p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(opt.arena.params, opt.arena.grad_views))
        lo, go, _ = O.vit_loss_and_grads(sd, cfg, xb, yb, torch.tensor(1.3), dtype=torch.float64)
        assert abs(float(loss) - float(lo)) < 2e-2
        grad_check(m, {k: 2 * v for k, v in go.items()}, GRAD_RTOL, f"arena/B={B}")
    F.enable_direct_grads(False)


def test_graphed_fused_train_step_has_only_library_kernels_and_matches_eager():
    """GraphedTrainStep over the fused engine: same weights as eager steps (dropout off), launch budget per step."""
    res = []
    launches = None
    for graphed in (False, True):
        cfg = vit3d_b200.get_config(16, 512, 2, 256, 8, dropout_rate=0.0)
        m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16")
        m.load_state_dict(O.init_state_dict(cfg, seed=42))
        m.to(DEV).train()
        opt = FusedSGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-2)
        x = O.synth_volumes(4, seed=7).to(DEV)
        y = O.synth_labels(4).to(DEV)
        if graphed:
            step = GraphedTrainStep(m, opt, warmup=3)
            for _ in range(4):
                loss = step(x, y, 1.5)
            assert step.fused
            launches = step.launches_per_replay
            assert opt._steps == 7
        else:
            for _ in range(7):
                opt.zero_grad()
                loss = m(x, y, 1.5)
                loss.backward()
                opt.step()
        torch.cuda.synchronize()
        res.append((float(loss), {k: v.detach().clone() for k, v in m.state_dict().items()}))
        F._STATE["step_dev"] = None
        F.enable_direct_grads(False)
    (le, se), (lg, sg) = res
    assert abs(le - lg) < 2e-3 * max(1.0, abs(le)), (le, lg)
    for k in se:
        d = float((se[k] - sg[k]).abs().max())
        assert d <= 2e-3 * float(se[k].abs().max()) + 1e-6, (k, d)
    assert launches is not None and launches <= 17 * 2 + 22, launches


def test_graphed_inference_recaptures_after_weight_update():
    """a graph captured before an optimizer step / load_state_dict must not replay stale weight shadows."""
    cfg = vit3d_b200.get_config(16, 512, 2, 256, 8, dropout_rate=0.0)
    m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision="bf16")
    m.load_state_dict(O.init_state_dict(cfg, seed=42))
    m.to(DEV).eval()
    g = GraphedInference(m)
    x = O.synth_volumes(4, seed=3).to(DEV)
    first = g(x)[0].clone()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.05)                                   # bumps every parameter's version
    with torch.no_grad():
        eager = m(x)[0].clone()
    again = g(x)[0].clone()
    assert g.recaptures == 1
    assert torch.equal(again, eager) and not torch.equal(again, first)
    m.load_state_dict(O.init_state_dict(cfg, seed=43))
    with torch.no_grad():
        eager2 = m(x)[0].clone()
    assert torch.equal(g(x)[0], eager2)


def _gelu_fit(x):
    """Fitted tanh-form GELU the BF16 forward evaluates and its derivative (csrc/tc_epilogue.cuh gelu_and_grad_pair)."""
    c0, c1, c2 = 0.797458471, 0.0370503451, -3.58732362e-4
    x2 = x * x
    t = torch.tanh(x * (c0 + c1 * x2 + c2 * x2 * x2))
    return 0.5 * x * (1 + t), 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * (c0 + 3 * c1 * x2 + 5 * c2 * x2 * x2)


def test_refresh_shadows_all_kinds_and_ragged_shapes():
    """vit3d_refresh_shadows: one launch over a job table rewrites bf16 / transposed-bf16 / fp16 / tf32-rounded / fp32
    copies of fp32 masters (64 x 64 tiles, vectorised where rows allow) - full tiles, ragged edges, row strides wider
    than the matrix (the packed q|k|v shadow), unaligned destinations - and bumps the device step counter."""
    import numpy as np
    gen = torch.Generator(device=DEV).manual_seed(7)
    shapes = [(256, 3072, 0), (3072, 256, 1), (100, 70, 0), (100, 70, 1), (65, 33, 2), (256, 5120, 3), (1, 256, 4),
              (130, 257, 3), (256, 256, 0), (77, 128, 1)]
    srcs, dsts, refs, rows_ = [], [], [], []
    arr = np.zeros(len(shapes) + 2, dtype=fused_train._JOB_DTYPE)
    t0 = 0
    T = _lib.SHADOW_TILE

    def add(i, src, dst_ptr, rows, cols, ld, kind):
        nonlocal t0
        tc = (cols + T - 1) // T
        arr[i] = (src.data_ptr(), dst_ptr, rows, cols, ld, kind, t0, tc)
        t0 += tc * ((rows + T - 1) // T)

    for i, (rows, cols, kind) in enumerate(shapes):
        src = torch.randn(rows, cols, device=DEV, generator=gen)
        if kind == 1:
            ld = rows + (8 if rows % 8 == 0 else 3)
            dst = torch.full((cols, ld), 7.0, device=DEV, dtype=torch.bfloat16)
            ref = src.t().to(torch.bfloat16)
            view = dst[:, :rows]
        else:
            ld = cols + (16 if cols % 8 == 0 else 5)
            dt = {0: torch.bfloat16, 2: torch.float16, 3: torch.float32, 4: torch.float32}[kind]
            dst = torch.full((rows, ld), 7.0, device=DEV, dtype=dt)
            ref = src.to(dt)
            if kind == 3:
                b = src.view(torch.int32)
                ref = (((b + 0x1000) & ~0x1FFF).view(torch.float32))            # round to nearest, ties away (cvt.rna)
            view = dst[:, :cols]
        add(i, src, dst.data_ptr(), rows, cols, ld, kind)
        srcs.append(src); dsts.append((dst, view, ld)); refs.append(ref)
    # the packed q|k|v shadow: two jobs into row blocks of ONE [512, 256] buffer
    packed = torch.full((512, 256), 7.0, device=DEV, dtype=torch.bfloat16)
    for j in range(2):
        src = torch.randn(256, 256, device=DEV, generator=gen)
        add(len(shapes) + j, src, packed.data_ptr() + j * 256 * 256 * 2, 256, 256, 256, 0)
        srcs.append(src)
    jobs = torch.from_numpy(arr.view(np.uint8).copy()).to(DEV)
    step = torch.tensor([41], device=DEV, dtype=torch.int32)
    _lib.call("vit3d_refresh_shadows", jobs.data_ptr(), len(arr), t0, step.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(step) == 42
    for (dst, view, ld), ref, (rows, cols, kind) in zip(dsts, refs, shapes):
        assert torch.equal(view, ref), (rows, cols, kind)
        pad = dst[:, (rows if kind == 1 else cols):]
        assert bool((pad == 7.0).all()), ("padding columns were written", rows, cols, kind)
    assert torch.equal(packed[:256], srcs[-2].to(torch.bfloat16)) and torch.equal(packed[256:], srcs[-1].to(torch.bfloat16))


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("M,K,drop", [(16640, 3072, True), (256, 1024, False), (19500 + 76, 2048, True), (130, 3072, True), (1000, 256, True)])
def test_linear_residual_dropout_layernorm_training_forward(M, K, drop, pair):
    """vit3d_linear_res_train_fwd: y = residual + Dropout(x w^T + b), ln_out = LayerNorm(y) (+ mean / rstd), against fp32
    torch on the same bf16 operands - as independent CTAs (pair=0) and as clusters of two CTAs sharing the weight
    k-blocks by TMA multicast (pair=1; taken for even tile counts with K >= 1024: 130 tiles at conf-18 batch 256,
    153 -> no, 154 tiles = several per CTA on 148 SMs; the other shapes fall back to single CTAs)."""
    H = 256
    gen = torch.Generator(device=DEV).manual_seed(M + K)
    st = torch.cuda.current_stream().cuda_stream
    x = (torch.randn(M, K, device=DEV, generator=gen) * 0.5).bfloat16()
    w = (torch.randn(H, K, device=DEV, generator=gen) / K ** 0.5).bfloat16()
    b = torch.randn(H, device=DEV, generator=gen) * 0.1
    res = torch.randn(M, H, device=DEV, generator=gen)
    gamma = 1 + 0.1 * torch.randn(H, device=DEV, generator=gen)
    beta = 0.1 * torch.randn(H, device=DEV, generator=gen)
    keep, bits, scale = torch.ones(M, H, device=DEV), None, 1.0
    if drop:
        keep = (torch.rand(M, H, device=DEV, generator=gen) > 0.1).float()
        bits = fused_train._pack_mask_bits(keep, DEV)
        scale = 1.0 / 0.9
    y = torch.empty(M, H, device=DEV)
    ln = torch.empty(M, H, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    _lib.lib().vit3d_set_tuning(8, pair)
    try:
        _lib.call("vit3d_linear_res_train_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), res.data_ptr(), y.data_ptr(),
                  None if bits is None else bits.data_ptr(), scale, gamma.data_ptr(), beta.data_ptr(), 1e-6, ln.data_ptr(),
                  mean.data_ptr(), rstd.data_ptr(), M, H, K, st)
        torch.cuda.synchronize()
    finally:
        _lib.lib().vit3d_set_tuning(8, 0)
    ref = res + (x.float() @ w.float().t() + b) * keep * scale
    assert float((y - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    mu, var = ref.mean(-1), ref.var(-1, unbiased=False)
    assert float((mean - mu).abs().max()) < 2e-3
    assert float((rstd - (var + 1e-6).rsqrt()).abs().max()) < 2e-3 * float((var + 1e-6).rsqrt().max())
    ref_ln = torch.nn.functional.layer_norm(ref, (H,), gamma, beta, 1e-6)
    assert float((ln.float() - ref_ln).abs().max()) < 0.03 * float(ref_ln.abs().max())


@pytest.mark.parametrize("M,d,drop", [(130, 256, True), (1000, 512, False), (19500, 512, True), (65, 3072, True)])
def test_fc1_train_forward_and_fused_mlp_backward_kernels(M, d, drop):
    """vit3d_fc1_train_fwd (act = Dropout(gelu(.)), dact = gelu'(.) x keep / (1-p)) and vit3d_mlp_bwd (dgrad fc2 -> x dact
    -> dgrad fc1 + fc1 bias gradient) against fp32 torch on the same bf16 operands; M = 19500 gives 153 row tiles on 148
    SMs (several tiles per CTA), M = 65 a single ragged tile."""
    H = 256
    gen = torch.Generator(device=DEV).manual_seed(M + d)
    st = torch.cuda.current_stream().cuda_stream
    xn = (torch.randn(M, H, device=DEV, generator=gen)).bfloat16()
    w1 = (torch.randn(d, H, device=DEV, generator=gen) / 16).bfloat16()
    b1 = torch.randn(d, device=DEV, generator=gen) * 0.1
    bits = None
    keep = torch.ones(M, d, device=DEV)
    scale = 1.0
    if drop:
        keep = (torch.rand(M, d, device=DEV, generator=gen) > 0.1).float()
        bits = fused_train._pack_mask_bits(keep, DEV)
        scale = 1.0 / 0.9
    dact = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    act = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    _lib.call("vit3d_fc1_train_fwd", xn.data_ptr(), w1.data_ptr(), b1.data_ptr(), dact.data_ptr(), act.data_ptr(),
              None if bits is None else bits.data_ptr(), scale, M, d, H, st)
    pre = xn.float() @ w1.float().t() + b1
    y, dy = _gelu_fit(pre)
    assert float((act.float() - y * keep * scale).abs().max()) < 0.03 * float(y.abs().max())
    assert float((dact.float() - dy * keep * scale).abs().max()) < 0.02

    gy = (torch.randn(M, H, device=DEV, generator=gen) * 0.05).bfloat16()
    w2_t = (torch.randn(d, H, device=DEV, generator=gen) / 16).bfloat16()          # W2^T
    w1_t = w1.t().contiguous()                                                     # W1^T
    dh = torch.empty(M, d, device=DEV, dtype=torch.bfloat16)
    dxn = torch.empty(M, H, device=DEV)
    db1 = torch.zeros(d, device=DEV)
    _lib.call("vit3d_mlp_bwd", gy.data_ptr(), w2_t.data_ptr(), w1_t.data_ptr(), dact.data_ptr(), dh.data_ptr(), dxn.data_ptr(),
              db1.data_ptr(), M, H, d, st)
    dh_ref = (gy.float() @ w2_t.float().t()) * dact.float()
    err = float((dh.float() - dh_ref).norm() / dh_ref.norm())
    assert err < 5e-3, err
    dx_ref = dh.float() @ w1_t.float().t()                    # from the kernel's own (bf16) dh: isolates the second GEMM
    errx = float((dxn - dx_ref).norm() / dx_ref.norm())
    assert errx < 2e-3, errx
    errb = float((db1 - dh_ref.sum(0)).norm() / dh_ref.sum(0).norm())
    assert errb < 5e-3, errb
    # the unfused element-wise stage gives the same dh and bias gradient
    da = (gy.float() @ w2_t.float().t()).bfloat16()
    dh2 = torch.empty_like(dh)
    db2 = torch.zeros(d, device=DEV)
    _lib.call("vit3d_mul_colsum_bwd", da.data_ptr(), dact.data_ptr(), dh2.data_ptr(), db2.data_ptr(), M, d, st)
    assert float((dh2.float() - dh_ref).norm() / dh_ref.norm()) < 1e-2
    assert float((db2 - dh_ref.sum(0)).norm() / dh_ref.sum(0).norm()) < 1e-2


def test_fused_mlp_backward_equals_three_pass_chain_in_the_training_step():
    cfg, sd, m, x, y, w = build("conf18")
    m.train()
    torch.manual_seed(7)
    grads = []
    for fused in (True, False):
        F._STATE["fused_mlp_bwd"] = fused
        F._STATE["step"] = 100                      # same dropout masks in both runs
        m.zero_grad(set_to_none=True)
        loss = m(x.to(DEV), y.to(DEV), w)
        loss.backward()
        grads.append({k: p.grad.detach().clone() for k, p in m.named_parameters()})
    F._STATE["fused_mlp_bwd"] = True
    gmax = max(float(v.norm()) for v in grads[0].values())
    for k in grads[0]:
        err = float((grads[0][k] - grads[1][k]).norm())
        assert err <= 0.03 * float(grads[1][k].norm()) + 2e-4 * gmax, (k, err)


@pytest.mark.parametrize("red", [2, 1, 0])
@pytest.mark.parametrize("M,N,K,seg", [(16640, 256, 3072, 0), (16640, 3072, 256, 0), (1000, 768, 256, 256), (65, 256, 256, 0),
                                       (777, 192, 320, 0), (300, 640, 64, 256)])
def test_wgrad_partial_tiles_plus_reduce_equal_atomic_wgrad(M, N, K, seg, red):
    """vit3d_wgrad_partial + vit3d_wgrad_reduce (no atomics) == vit3d_wgrad (red=1: tiles added by TMA reduce boxes, red=2: the same from 256-row tiles,
    red=0: fp32 vector atomics) == fp32 torch, including the packed q|k|v product whose row segments land in three
    buffers (the last one ragged), output rows that do not fill a tile, accumulating (+=) into existing gradients."""
    import ctypes as C
    _lib.lib().vit3d_set_tuning(6, red)
    gen = torch.Generator(device=DEV).manual_seed(M + N + K)
    dy = (torch.randn(M, N, device=DEV, generator=gen) * 0.1).bfloat16()
    x = (torch.randn(M, K, device=DEV, generator=gen) * 0.5).bfloat16()
    ref = dy.float().t() @ x.float()
    st = torch.cuda.current_stream().cuda_stream
    nseg = 3 if seg else 1
    rows = [min(seg, N - i * seg) for i in range(nseg)] if seg else [N]
    base = [torch.full((r, K), 0.25, device=DEV) for r in rows]
    atom = [b.clone() for b in base]
    part = [b.clone() for b in base]
    pa = [t.data_ptr() for t in atom] + [None] * (3 - nseg)
    try:
        _lib.call("vit3d_wgrad", dy.data_ptr(), x.data_ptr(), pa[0], pa[1], pa[2], seg, M, N, K, st)
    finally:
        _lib.lib().vit3d_set_tuning(6, 1)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    got_a = torch.cat(atom) - 0.25
    assert float((got_a - ref).abs().max()) <= 2e-3 * scale + 1e-4
    if N % 128:          # the partial-tile variant needs whole 128-row output tiles
        return
    ws = torch.empty(_lib.lib().vit3d_wgrad_ws_bytes(M, N, K) // 4, device=DEV)
    bn, sp = C.c_int(0), C.c_int(0)
    _lib.call("vit3d_wgrad_partial", dy.data_ptr(), x.data_ptr(), ws.data_ptr(), M, N, K, C.byref(bn), C.byref(sp), st)
    job = (fused_train._WgradJob * 1)()
    job[0].ws = ws.data_ptr()
    for i in range(nseg):
        job[0].dst[i] = part[i].data_ptr()
    job[0].seg_rows, job[0].rows, job[0].cols, job[0].bn, job[0].splits, job[0].tiles_n = seg, N, K, bn.value, sp.value, K // bn.value
    _lib.call("vit3d_wgrad_reduce", C.cast(job, C.c_void_p), 1, st)
    torch.cuda.synchronize()
    got_p = torch.cat(part) - 0.25
    assert float((got_p - ref).abs().max()) <= 2e-3 * scale + 1e-4
    assert float((got_a - got_p).abs().max()) <= 1e-3 * scale + 1e-4

"""CPU oracle for the 3D-ViT stacking-ensemble forward/backward path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and only as the checker / the CPU arm.

It is a *functional restatement* (plain torch CPU ops on explicit tensors, no
``nn.Module`` forward code) of the arithmetic of the reference file
``models/modeling.py`` (evapachetti/3d_vit_ensemble):

  * ``embeddings``        <- Embeddings.forward            modeling.py:162-175
  * ``attention``         <- Attention.forward             modeling.py:78-99
  * ``mlp``               <- Mlp.forward                   modeling.py:118-124
  * ``block``             <- Block.forward                 modeling.py:187-197
  * ``encoder``           <- Encoder.forward               modeling.py:247-254
  * ``vit_forward``       <- VisionTransformer.forward     modeling.py:279-288
  * ``ensemble_forward``  <- TransformerEnsemble.forward   modeling.py:353-356
  * ``init_state_dict``   <- the constructors' parameter-creation order
                             (modeling.py:56-71,103-116,130-160,179-185,238-245,270-277)

Backward is obtained from torch autograd over this restatement (fp32 or fp64).

Parity pin: the reference ships no tests / golden vectors, so the
pin is the reference itself: ``oracle/gen_golden.py`` imports the unmodified
reference from /root/reference, runs it on seeded inputs and commits the results
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors (``-m "not gpu"``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

Z_SIZE = 5  # slices per volume, modeling.py:134


class Cfg(dict):
    """Attribute+item access config object; stands in for ml_collections.ConfigDict
    (absent in this image).  Same attribute contract as tools.py:84-97."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def get_config(ps: int, dim: int, n: int, hs: int, nh: int, dropout: float = 0.1) -> Cfg:
    """tools.py:84-97 (get_config) with the same field names."""
    c = Cfg()
    c.patches = Cfg(size=(ps, ps, Z_SIZE))
    c.hidden_size = hs
    c.transformer = Cfg(mlp_dim=dim, num_heads=nh, num_layers=n,
                        attention_dropout_rate=0.0, dropout_rate=dropout)
    c.classifier = "token"
    c.representation_size = None
    return c


# README.md:24-44 table read as head-dim D x heads k = hidden 256
def north_star_config(conf: int, dropout: float = 0.1) -> Cfg:
    assert 1 <= conf <= 18
    i = conf - 1
    d = (2048, 3072)[i // 9]
    L = (4, 6, 8)[(i % 9) // 3]
    k = (4, 8, 16)[i % 3]
    return get_config(16, d, L, 256, k, dropout)


def as_shipped_config(conf: int, dropout: float = 0.1) -> Cfg:
    """What tools.parameters_config(conf) really returns (tools.py:60-80): the dict key
    is overwritten by every inner-loop combination, so each conf gets the LAST one."""
    if 1 <= conf <= 18:
        return get_config(16, 3072, 8, 16, 16, dropout)
    if 19 <= conf <= 26:
        return get_config(8, 2204, 6, 8, 8, dropout)
    raise KeyError(conf)


def n_patches(cfg, img_size: int) -> int:
    ps = cfg.patches["size"]
    return (img_size // ps[0]) * (img_size // ps[1]) * (Z_SIZE // ps[2])


# --------------------------------------------------------------------------- init
def init_state_dict(cfg, img_size: int = 128, num_classes: int = 1, seed: Optional[int] = 42,
                    randomize_tokens: bool = True) -> Dict[str, torch.Tensor]:
    """Create parameters in the reference constructors' order with the same torch
    initialisers, so the same seed gives the same weights as the reference modules.
    ``randomize_tokens``: cls/pos-emb are zeros in the reference (modeling.py:157-158);
    re-draw them N(0, 0.02) *after* everything else so the fused
    add/concat is exercised."""
    if seed is not None:
        torch.manual_seed(seed)
    H = cfg.hidden_size
    d = cfg.transformer["mlp_dim"]
    L = cfg.transformer["num_layers"]
    ps = tuple(cfg.patches["size"])
    P = n_patches(cfg, img_size)
    sd: Dict[str, torch.Tensor] = {}
    # Embeddings.__init__ : Conv3d, pos zeros, cls zeros            modeling.py:153-158
    conv = nn.Conv3d(1, H, kernel_size=ps, stride=ps)
    sd["transformer.embeddings.position_embeddings"] = torch.zeros(1, P + 1, H)
    sd["transformer.embeddings.cls_token"] = torch.zeros(1, 1, H)
    sd["transformer.embeddings.patch_embeddings.weight"] = conv.weight.detach().clone()
    sd["transformer.embeddings.patch_embeddings.bias"] = conv.bias.detach().clone()
    # Encoder.__init__ : encoder_norm first, then L freshly built Blocks  modeling.py:241-245
    enc_norm = nn.LayerNorm(H, eps=1e-6)
    layers = []
    for _ in range(L):
        # Block.__init__ order: attention_norm, ffn_norm, Mlp, Attention   modeling.py:182-185
        an = nn.LayerNorm(H, eps=1e-6)
        fn = nn.LayerNorm(H, eps=1e-6)
        fc1 = nn.Linear(H, d)
        fc2 = nn.Linear(d, H)
        nn.init.xavier_uniform_(fc1.weight)          # Mlp._init_weights modeling.py:112-116
        nn.init.xavier_uniform_(fc2.weight)
        nn.init.normal_(fc1.bias, std=1e-6)
        nn.init.normal_(fc2.bias, std=1e-6)
        q = nn.Linear(H, H)
        k = nn.Linear(H, H)
        v = nn.Linear(H, H)
        o = nn.Linear(H, H)
        layers.append((an, fn, fc1, fc2, q, k, v, o))
    for i, (an, fn, fc1, fc2, q, k, v, o) in enumerate(layers):
        p = f"transformer.encoder.layer.{i}."
        for name, m in (("attention_norm", an), ("ffn_norm", fn), ("ffn.fc1", fc1), ("ffn.fc2", fc2),
                        ("attn.query", q), ("attn.key", k), ("attn.value", v), ("attn.out", o)):
            sd[p + name + ".weight"] = m.weight.detach().clone()
            sd[p + name + ".bias"] = m.bias.detach().clone()
    sd["transformer.encoder.encoder_norm.weight"] = enc_norm.weight.detach().clone()
    sd["transformer.encoder.encoder_norm.bias"] = enc_norm.bias.detach().clone()
    head = nn.Linear(H, num_classes)                                   # modeling.py:277
    sd["head.weight"] = head.weight.detach().clone()
    sd["head.bias"] = head.bias.detach().clone()
    if randomize_tokens:
        sd["transformer.embeddings.cls_token"] = torch.randn(1, 1, H) * 0.02
        sd["transformer.embeddings.position_embeddings"] = torch.randn(1, P + 1, H) * 0.02
    return sd


def synth_volumes(B: int, seed: int = 42, kind: str = "img", img_size: int = 128) -> torch.Tensor:
    """Synthetic single-channel T2w-like volumes (B,1,img,img,5) fp32.
    'img': clamp(round(66+45*N(0,1)),0,255) minus its mean; 'unit': N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(B, 1, img_size, img_size, Z_SIZE, generator=g)
    if kind == "unit":
        return r
    u8 = torch.clamp(torch.round(66.0 + 45.0 * r), 0, 255)
    return (u8 - u8.mean()).float()


def synth_volumes_u8(B: int, seed: int = 42, img_size: int = 128):
    """The same 'img' volumes as synth_volumes, before the mean subtraction: (uint8 tensor, mean).
    synth_volumes(B, seed) == u8.float() - mean exactly."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(B, 1, img_size, img_size, Z_SIZE, generator=g)
    u8 = torch.clamp(torch.round(66.0 + 45.0 * r), 0, 255)
    return u8.to(torch.uint8), float(u8.mean())


def synth_labels(B: int, seed: int = 42) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + 1)
    y = torch.randint(0, 2, (B,), generator=g).float()
    if B >= 2:  # make sure both classes exist so the balanced weight is finite
        y[0], y[1] = 0.0, 1.0
    return y


def balanced_pos_weight(y: torch.Tensor) -> Optional[torch.Tensor]:
    """TEST-FIXTURE weight: the ratio n_neg/n_pos.  tests/golden/*.npz were generated with this value fed
    to the unmodified reference as `weights`, so it stays; it is NOT what the scripts compute per batch -
    that is `sklearn_pos_weight` below (the two agree only on balanced batches)."""
    n_pos = float(y.sum())
    n_neg = float(y.numel()) - n_pos
    if n_pos == 0 or n_neg == 0:
        return None
    return torch.tensor(n_neg / n_pos, dtype=torch.float64)


def sklearn_pos_weight(y: torch.Tensor) -> torch.Tensor:
    """What train_baseline_cv.py:168-169 passes as pos_weight: compute_class_weight('balanced', classes=
    unique(y), y) = n / (n_classes_present * bincount); entry [1] (= n / (2 n_pos)) when both classes are
    present, else entry [0] = 1.0."""
    n = float(y.numel())
    n_pos = float(y.sum())
    if n_pos == 0 or n_pos == n:
        return torch.tensor(1.0, dtype=torch.float64)
    return torch.tensor(n / (2.0 * n_pos), dtype=torch.float64)


# --------------------------------------------------------------------------- forward
def patch_gather(x: torch.Tensor, ps: Sequence[int]) -> torch.Tensor:
    """im2col of a stride==kernel Conv3d as a pure permutation: (B,1,X,Y,Z) ->
    (B, P, ps0*ps1*ps2) with token order p = (px*ny + py)*nz + pz (flatten(2) of the
    conv output, modeling.py:168-170) and K order (i*ps1 + j)*ps2 + z (= weight.view)."""
    B, C, X, Y, Z = x.shape
    assert C == 1
    nx, ny, nz = X // ps[0], Y // ps[1], Z // ps[2]
    x = x[:, 0, :nx * ps[0], :ny * ps[1], :nz * ps[2]]
    x = x.reshape(B, nx, ps[0], ny, ps[1], nz, ps[2])
    x = x.permute(0, 1, 3, 5, 2, 4, 6)
    return x.reshape(B, nx * ny * nz, ps[0] * ps[1] * ps[2])


def _dropout(x, mask, p):
    if mask is None:
        return x
    return x * mask.to(x.dtype) / (1.0 - p)


def embeddings(sd, cfg, x, pre="transformer.embeddings.", mask=None):
    H = cfg.hidden_size
    ps = tuple(cfg.patches["size"])
    w = sd[pre + "patch_embeddings.weight"].reshape(H, -1)
    pt = patch_gather(x, ps) @ w.t() + sd[pre + "patch_embeddings.bias"]
    B = x.shape[0]
    cls = sd[pre + "cls_token"].expand(B, -1, -1)
    tok = torch.cat((cls, pt), dim=1) + sd[pre + "position_embeddings"]
    return _dropout(tok, mask, cfg.transformer["dropout_rate"])


def layer_norm(x, w, b, eps=1e-6):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)      # biased, as nn.LayerNorm
    return (x - mu) / torch.sqrt(var + eps) * w + b


def attention(sd, cfg, x, pre):
    B, S, H = x.shape
    k = cfg.transformer["num_heads"]
    D = int(H / k)
    A = k * D

    def heads(t):
        return t.view(B, S, k, D).permute(0, 2, 1, 3)

    q = heads(x @ sd[pre + "query.weight"].t() + sd[pre + "query.bias"])
    kk = heads(x @ sd[pre + "key.weight"].t() + sd[pre + "key.bias"])
    v = heads(x @ sd[pre + "value.weight"].t() + sd[pre + "value.bias"])
    scores = (q @ kk.transpose(-1, -2)) / math.sqrt(D)
    probs = torch.softmax(scores, dim=-1)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, S, A)
    return ctx @ sd[pre + "out.weight"].t() + sd[pre + "out.bias"], probs


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def mlp(sd, cfg, x, pre, m1=None, m2=None):
    p = cfg.transformer["dropout_rate"]
    h = gelu_erf(x @ sd[pre + "fc1.weight"].t() + sd[pre + "fc1.bias"])
    h = _dropout(h, m1, p)
    y = h @ sd[pre + "fc2.weight"].t() + sd[pre + "fc2.bias"]
    return _dropout(y, m2, p)


def block(sd, cfg, x, pre, m1=None, m2=None):
    a, probs = attention(sd, cfg, layer_norm(x, sd[pre + "attention_norm.weight"],
                                             sd[pre + "attention_norm.bias"]), pre + "attn.")
    x = x + a
    y = mlp(sd, cfg, layer_norm(x, sd[pre + "ffn_norm.weight"], sd[pre + "ffn_norm.bias"]),
            pre + "ffn.", m1, m2)
    return x + y, probs


def encoder(sd, cfg, x, pre="transformer.encoder.", masks=None):
    probs_all = []
    hidden = []
    for i in range(cfg.transformer["num_layers"]):
        m1 = m2 = None
        if masks is not None:
            m1, m2 = masks.get(("fc1", i)), masks.get(("fc2", i))
        x, probs = block(sd, cfg, x, f"{pre}layer.{i}.", m1, m2)
        probs_all.append(probs)
        hidden.append(x)
    enc = layer_norm(x, sd[pre + "encoder_norm.weight"], sd[pre + "encoder_norm.bias"])
    return enc, probs_all, hidden


def bce_with_logits(logits, labels, pos_weight=None):
    """BCEWithLogitsLoss(pos_weight)(logits.view(-1,1), labels.view(-1,1)), mean.
    modeling.py:283-286."""
    z = logits.reshape(-1)
    y = labels.reshape(-1).to(z.dtype)
    pw = 1.0 if pos_weight is None else pos_weight.to(z.dtype)
    loss = -(pw * y * F.logsigmoid(z) + (1.0 - y) * F.logsigmoid(-z))
    return loss.mean()


def vit_forward(sd, cfg, x, labels=None, weights=None, masks=None, prefix="", want_hidden=False):
    """VisionTransformer.forward, modeling.py:279-288.  ``masks`` (optional) injects
    dropout keep-masks: {'emb': (B,S,H), ('fc1',l): (B,S,d), ('fc2',l): (B,S,H)}."""
    sub = sd if not prefix else {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    tok = embeddings(sub, cfg, x, mask=None if masks is None else masks.get("emb"))
    enc, probs, hidden = encoder(sub, cfg, tok, masks=masks)
    logits = enc[:, 0] @ sub["head.weight"].t() + sub["head.bias"]
    if labels is not None:
        return bce_with_logits(logits, labels, weights)
    if want_hidden:
        return logits, probs, enc, [tok] + hidden
    return logits, probs, enc


def ensemble_forward(sd, cfgs, x):
    """TransformerEnsemble.forward (in_features == num_classes == 1), modeling.py:353-356."""
    outs = [vit_forward(sd, c, x, prefix=f"transformers.{j}.")[0] for j, c in enumerate(cfgs)]
    cat = torch.cat(outs, dim=1)
    return torch.sigmoid(cat @ sd["classifier.weight"].t() + sd["classifier.bias"])


def ensemble_state_dict(member_sds, seed: Optional[int] = 7):
    if seed is not None:
        torch.manual_seed(seed)
    sd = {}
    for j, m in enumerate(member_sds):
        for k, v in m.items():
            sd[f"transformers.{j}.{k}"] = v
    clf = nn.Linear(len(member_sds), 1)
    sd["classifier.weight"] = clf.weight.detach().clone()
    sd["classifier.bias"] = clf.bias.detach().clone()
    return sd


# --------------------------------------------------------------------------- grads
def to_dtype(sd, dtype):
    return {k: v.detach().to(dtype).clone() for k, v in sd.items()}


def vit_loss_and_grads(sd, cfg, x, labels, weights=None, masks=None, dtype=torch.float32):
    """Returns (loss, {param: grad}, dx) from autograd over the restatement."""
    p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    mm = None if masks is None else {k: v.to(dtype) for k, v in masks.items()}
    loss = vit_forward(p, cfg, xx, labels.to(dtype), weights, mm)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return loss.detach(), grads, xx.grad


def fwd_flops_per_volume(cfg, img_size=128) -> float:
    """Algorithmic forward FLOPs per volume (2 per MAC), SURVEY.md §8d formula."""
    H = cfg.hidden_size
    d = cfg.transformer["mlp_dim"]
    L = cfg.transformer["num_layers"]
    ps = cfg.patches["size"]
    P = n_patches(cfg, img_size)
    S = P + 1
    Kp = ps[0] * ps[1] * ps[2]
    return 2.0 * P * Kp * H + L * (6.0 * S * H * H + 4.0 * S * S * H + 2.0 * S * H * H + 4.0 * S * H * d) + 2.0 * H

"""Generate golden vectors from the UNMODIFIED reference (test infrastructure).

Runs only in the build container, where the reference is mounted read-only at
/root/reference (it does not exist on the GPU box).  It imports
``models.modeling`` from there, drives ``VisionTransformer`` /
``TransformerEnsemble`` on seeded synthetic volumes and writes small ``.npz``
fixtures to ``tests/golden/``.  ``tests/test_oracle_golden.py`` then pins
``oracle/vit3d_oracle.py`` against them, and the ``-m gpu`` tests pin the CUDA path
against both.

    python oracle/gen_golden.py            # regenerates every fixture
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("VIT3D_REFERENCE", "/root/reference")

from oracle import vit3d_oracle as O  # noqa: E402

CASES = {
    # name: (cfg-args (ps, mlp, L, hidden, heads), B)
    "tiny": ((16, 64, 2, 32, 4), 3),
    "shipped": ((16, 3072, 8, 16, 16), 2),      # what every conf 1..18 really builds (tools.py:60-80)
    "shipped_p8": ((8, 2204, 6, 8, 8), 1),      # conf 19..26 as shipped: 256 patches + cls
    "conf5": ((16, 2048, 6, 256, 8), 2),
    "conf9": ((16, 2048, 8, 256, 16), 2),
    "conf11": ((16, 3072, 4, 256, 8), 2),
    "conf18": ((16, 3072, 8, 256, 16), 2),
    "conf1": ((16, 2048, 4, 256, 4), 2),
}


def ref_modules():
    sys.path.insert(0, REF)
    from models import modeling  # the reference, unmodified
    return modeling


def stats(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().reshape(-1)
    head = torch.zeros(8, dtype=torch.float64)
    n = min(8, t.numel())
    head[:n] = t[:n]
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) / t.numel()
    return torch.cat([torch.stack([t.sum(), t.norm(), t.abs().max(), (t * w).sum()]), head]).numpy()


def build_ref(M, cfg, seed=42, randomize_tokens=True, vis=True):
    torch.manual_seed(seed)
    m = M.VisionTransformer(cfg, 128, zero_head=True, num_classes=1, vis=vis)
    if randomize_tokens:
        with torch.no_grad():
            H = cfg.hidden_size
            P = O.n_patches(cfg, 128)
            m.transformer.embeddings.cls_token.copy_(torch.randn(1, 1, H) * 0.02)
            m.transformer.embeddings.position_embeddings.copy_(torch.randn(1, P + 1, H) * 0.02)
    return m


def run_case(M, name, args, B, out):
    cfg = O.get_config(*args)
    m = build_ref(M, cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = O.synth_volumes(B, seed=42, kind="img")
    y = O.synth_labels(B)
    w = O.balanced_pos_weight(y)
    g = {"x_stats": stats(x), "labels": y.numpy()}
    g["pos_weight"] = np.array(-1.0 if w is None else float(w))
    for k, v in sd.items():
        g["sd_stats/" + k] = stats(v)
    # ---- eval forward (modeling.py:287-288)
    m.eval()
    with torch.no_grad():
        logits, probs, enc = m(x)
    g["logits"] = logits.numpy()
    g["cls_feature"] = enc[:, 0].numpy()
    g["enc_stats"] = stats(enc)
    for i, p in enumerate(probs):
        g[f"probs_stats/{i}"] = stats(p)
    if name == "tiny":
        g["enc"] = enc.numpy()
        g["probs0"] = probs[0].numpy()
    # unit-normal inputs too (numerics stress)
    xu = O.synth_volumes(B, seed=43, kind="unit")
    with torch.no_grad():
        g["logits_unit"] = m(xu)[0].numpy()
    # ---- loss + grads in eval mode (dropout off; SURVEY.md §7 'dropout parity')
    m.zero_grad()
    wt = None if w is None else w
    loss = m(x, y, wt)
    loss.backward()
    g["loss_eval"] = np.array(float(loss))
    for k, p in m.named_parameters():
        g["grad_eval_stats/" + k] = stats(p.grad)
        if name == "tiny":
            g["grad_eval/" + k] = p.grad.numpy().copy()
    # ---- loss + grads in train mode with the dropout masks the reference drew
    m.train()
    m.zero_grad()
    masks = {}
    hooks = []

    def mk(key, p):
        def hook(mod, inp, outp):
            i = inp[0]
            assert int((i == 0).sum()) == 0 or p == 0.0
            masks[key] = (outp != 0).detach()
        return hook

    p_drop = cfg.transformer["dropout_rate"]
    hooks.append(m.transformer.embeddings.dropout.register_forward_hook(mk("emb", p_drop)))
    for i, blk in enumerate(m.transformer.encoder.layer):
        # Mlp.forward calls self.dropout twice (after gelu, after fc2) modeling.py:121-123
        calls = {"n": 0}

        def hook2(mod, inp, outp, i=i, calls=calls):
            key = ("fc1", i) if calls["n"] % 2 == 0 else ("fc2", i)
            calls["n"] += 1
            masks[key] = (outp != 0).detach()
        hooks.append(blk.ffn.dropout.register_forward_hook(hook2))
    torch.manual_seed(1234)
    loss_t = m(x, y, wt)
    loss_t.backward()
    for h in hooks:
        h.remove()
    g["loss_train"] = np.array(float(loss_t))
    for k, p in m.named_parameters():
        g["grad_train_stats/" + k] = stats(p.grad)
    for key, mask in masks.items():
        kn = key if isinstance(key, str) else f"{key[0]}_{key[1]}"
        g["mask/" + kn] = np.packbits(mask.numpy().reshape(-1))
        g["mask_shape/" + kn] = np.array(mask.shape)
    np.savez_compressed(os.path.join(out, f"{name}.npz"), **g)
    # cross-check the restatement right here (fail generation if it is off)
    lo, po, eo = O.vit_forward(sd, cfg, x)
    err = float((lo - logits).abs().max())
    assert err < 1e-4, (name, err)
    print(f"{name}: B={B} logits={logits.flatten().tolist()} loss_eval={float(loss):.6f} "
          f"loss_train={float(loss_t):.6f} oracle_err={err:.2e}")
    return m, sd


def run_init_parity(M, out):
    """Same seed => same initial weights as the reference constructors."""
    g = {}
    for name, (args, _) in CASES.items():
        cfg = O.get_config(*args)
        torch.manual_seed(42)
        m = M.VisionTransformer(cfg, 128, zero_head=True, num_classes=1)
        for k, v in m.state_dict().items():
            g[f"{name}/{k}"] = stats(v)
    np.savez_compressed(os.path.join(out, "init_seed42.npz"), **g)


def run_ensemble(M, out):
    cfgs = [O.north_star_config(c) for c in (5, 9, 11)]
    members = []
    for j, c in enumerate(cfgs):
        members.append(build_ref(M, c, seed=42 + j))
    torch.manual_seed(7)
    ens = M.TransformerEnsemble(*members, in_features=1)
    ens.eval()
    x = O.synth_volumes(3, seed=42, kind="img")
    with torch.no_grad():
        outp = ens(x)
        member_logits = torch.cat([t(x)[0] for t in ens.transformers], dim=1)
    g = {"out": outp.numpy(), "member_logits": member_logits.numpy(),
         "classifier.weight": ens.classifier.weight.detach().numpy(),
         "classifier.bias": ens.classifier.bias.detach().numpy(),
         "n_keys": np.array(len(ens.state_dict()))}
    # ensemble training step gradient (train_ensemble_whole_dataset.py:115-123, with the
    # (B,1)-vs-(B,) label shape defect fixed by unsqueeze), eval-mode dropout
    y = O.synth_labels(3)
    ens.zero_grad()
    o = ens(x)
    loss = torch.nn.BCELoss()(o, y.unsqueeze(1))
    loss.backward()
    g["loss"] = np.array(float(loss))
    g["grad_stats/classifier.weight"] = stats(ens.classifier.weight.grad)
    g["grad_stats/classifier.bias"] = stats(ens.classifier.bias.grad)
    for j in range(3):
        k = f"transformers.{j}.head.weight"
        g["grad_stats/" + k] = stats(dict(ens.named_parameters())[k].grad)
        k = f"transformers.{j}.transformer.embeddings.patch_embeddings.weight"
        g["grad_stats/" + k] = stats(dict(ens.named_parameters())[k].grad)
    np.savez_compressed(os.path.join(out, "ensemble_5_9_11.npz"), **g)
    sd = {k: v.detach().clone() for k, v in ens.state_dict().items()}
    err = float((O.ensemble_forward(sd, cfgs, x) - outp).abs().max())
    assert err < 1e-5, err
    print("ensemble 5+9+11:", outp.flatten().tolist(), "oracle_err=%.2e" % err)


def run_real_volumes(M, out):
    """A few real volumes through the reference's own ProstateDataset (create_dataset.py:14-69)
    as realistic fixtures (SURVEY.md §8c 'Fixtures')."""
    import pandas as pd
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        sys.path.insert(0, REF)
        from create_dataset import ProstateDataset
        ds = ProstateDataset(os.path.join(REF, "csv_files/fixed_split/test.csv"))
        vols, labs = [], []
        for i in range(min(4, len(ds))):
            v, l = ds[i][:2]
            vols.append(np.asarray(v))
            labs.append(int(l))
    except Exception as e:  # dataset class quirks must not block the other fixtures
        print("real-volume fixture skipped:", repr(e))
        os.chdir(cwd)
        return
    os.chdir(cwd)
    v = np.stack(vols)                                  # (n,128,128,5,1) float64 0..255
    u8 = v.astype(np.uint8)
    assert np.array_equal(u8.astype(v.dtype), v)
    x = torch.from_numpy(v).float().permute(0, 4, 1, 2, 3).contiguous()    # ToTensorDataset: (1,128,128,5)
    x = x - x.mean()
    cfg = O.north_star_config(5)
    m = build_ref(M, cfg)
    m.eval()
    with torch.no_grad():
        logits = m(x)[0]
    np.savez_compressed(os.path.join(out, "real_volumes.npz"), u8=u8, labels=np.array(labs),
                        mean=np.array(float(torch.from_numpy(v).float().mean())),
                        logits_conf5=logits.numpy())
    print("real volumes:", u8.shape, "logits", logits.flatten().tolist())


CKPT_CASES = [(16, 64, 2, 32, 4), (16, 96, 1, 32, 2), (16, 64, 3, 32, 8)]      # three small baselines (hidden 32)


def run_checkpoints(M, out):
    """N4: `.bin` files written by the UNMODIFIED reference classes the way train_baseline_cv.py:128-134 writes them
    (`torch.save(model.state_dict(), path)`), plus what the reference's own TransformerEnsemble makes of them.
    The GPU test loads these files through workflow.ensemble_from_checkpoints and must reproduce the outputs."""
    members, paths = [], []
    for j, args in enumerate(CKPT_CASES):
        m = build_ref(M, O.get_config(*args), seed=100 + j)
        path = os.path.join(out, f"ref_ckpt_member{j}.bin")
        model_to_save = m.module if hasattr(m, "module") else m          # train_baseline_cv.py:129
        torch.save(model_to_save.state_dict(), path)
        members.append(m)
        paths.append(path)
    torch.manual_seed(11)
    ens = M.TransformerEnsemble(*members, in_features=1)
    ens.eval()
    x = O.synth_volumes(3, seed=5, kind="img")
    with torch.no_grad():
        outp = ens(x)
        member_logits = torch.cat([t(x)[0] for t in ens.transformers], dim=1)
    np.savez_compressed(os.path.join(out, "ref_ckpt_ensemble.npz"), out=outp.numpy(), member_logits=member_logits.numpy(),
                        classifier_weight=ens.classifier.weight.detach().numpy(),
                        classifier_bias=ens.classifier.bias.detach().numpy(), cfg_args=np.array(CKPT_CASES))
    print("reference checkpoints:", [os.path.getsize(p) for p in paths], "bytes; ensemble out", outp.flatten().tolist())


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    torch.set_num_threads(8)
    M = ref_modules()
    if "--checkpoints" in sys.argv:          # only the checkpoint fixtures (the others are unchanged)
        run_checkpoints(M, out)
        return
    run_init_parity(M, out)
    for name, (args, B) in CASES.items():
        run_case(M, name, args, B, out)
    run_ensemble(M, out)
    run_real_volumes(M, out)
    run_checkpoints(M, out)


if __name__ == "__main__":
    main()

"""Worst logit error of the tensor-core modes against the reference-generated golden vectors (tests/golden)."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as T
worst = {}
for prec in ("tf32", "bf16"):
    w = 0.0
    for name in ["conf5", "conf9", "conf11", "conf18", "conf1", "shipped", "tiny"]:
        g = T.load_golden(name)
        cfg, sd, m, x, y, wt = T.build(name, prec)
        m.eval()
        with torch.no_grad():
            logits, probs, enc = m(x.to(T.DEV))
        err = float((logits.cpu() - torch.from_numpy(g["logits"])).abs().max())
        print(prec, name, f"{err:.2e}")
        w = max(w, err)
    worst[prec] = w
print(worst)

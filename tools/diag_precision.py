"""Per-layer error of a precision mode against the fp64 oracle (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit3d_b200
from oracle import vit3d_oracle as O
from vit3d_b200.models.modeling import VisionTransformer
from tests.helpers import CASES, case_setup, load_golden

for name in ["conf5", "conf9", "conf11", "conf18", "conf1"]:
    cfg, sd, x, y, w = case_setup(name)
    sd64 = {k: v.double() for k, v in sd.items()}
    lo, po, eo, hid = O.vit_forward(sd64, cfg, x.double(), want_hidden=True)
    for prec in sys.argv[1:] or ["tf32", "bf16"]:
        m = VisionTransformer(cfg, 128, zero_head=True, num_classes=1, precision=prec)
        m.load_state_dict(sd); m.to("cuda").eval()
        outs = []
        hooks = [m.transformer.embeddings.register_forward_hook(lambda mod, i, o: outs.append(o.detach().cpu().double()))]
        for blk in m.transformer.encoder.layer:
            hooks.append(blk.register_forward_hook(lambda mod, i, o: outs.append(o[0].detach().cpu().double())))
        with torch.no_grad():
            logits = m(x.cuda())[0].cpu().double()
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, hid)]
        print(name, prec, "logit err %.2e" % float((logits - lo).abs().max()), "rel hidden err per layer:",
              " ".join("%.1e" % e for e in errs))

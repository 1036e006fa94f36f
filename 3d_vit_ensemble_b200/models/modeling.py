"""`models.modeling` — the reference's import surface (train_baseline_cv.py:14, train_ensemble_cv.py:8).

Works both as `vit3d_b200.models.modeling` and, with the package directory on sys.path, as the
reference's own top-level `models.modeling`; either way the classes are the single implementation in
`3d_vit_ensemble_b200/vit.py` (no duplicate class objects)."""
try:
    from ..vit import (Attention, Block, Embeddings, Encoder, Mlp, Transformer, TransformerEnsemble,  # noqa: F401
                       VisionTransformer)
except ImportError:      # imported as top-level `models.modeling`: the parent package is not in scope
    import importlib
    import os
    import sys

    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    _vit = importlib.import_module("3d_vit_ensemble_b200.vit")
    Attention, Block, Embeddings, Encoder, Mlp = _vit.Attention, _vit.Block, _vit.Embeddings, _vit.Encoder, _vit.Mlp
    Transformer, TransformerEnsemble, VisionTransformer = _vit.Transformer, _vit.TransformerEnsemble, _vit.VisionTransformer

__all__ = ["Attention", "Block", "Embeddings", "Encoder", "Mlp", "Transformer", "TransformerEnsemble",
           "VisionTransformer"]

"""Model configuration objects with the attribute contract of the reference's
``tools.get_config`` (tools.py:84-97).  ``ml_collections`` is not required: any object with
attribute *and* item access works, and this small dict subclass is one."""
from __future__ import annotations


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def get_config(ps, dim, n, hs, nh, dropout_rate=0.1):
    """Same fields as tools.py:84-97: patch (ps,ps,5), mlp_dim, num_layers, hidden_size, num_heads."""
    c = ConfigDict()
    c.patches = ConfigDict(size=(ps, ps, 5))
    c.hidden_size = hs
    c.transformer = ConfigDict(mlp_dim=dim, num_heads=nh, num_layers=n, attention_dropout_rate=0.0,
                               dropout_rate=dropout_rate)
    c.classifier = "token"
    c.representation_size = None
    return c


def parameters_config(conf):
    """What the reference's ``tools.parameters_config`` RETURNS for each id (tools.py:60-80): its dict
    key is overwritten by every inner-loop combination, so ids 1..18 all give (16,3072,8,16,16) and
    19..26 give (8,2204,6,8,8).  Kept for drop-in behaviour; see ``north_star_config`` for the
    README table."""
    if 1 <= conf <= 18:
        return 16, 3072, 8, 16, 16
    if 19 <= conf <= 26:
        return 8, 2204, 6, 8, 8
    raise KeyError(f"Configuration {conf}")


def north_star_config(conf, dropout_rate=0.1):
    """README.md:24-44 read as head-dim D x heads k = hidden 256 (the intended 18 baselines):
    mlp {2048,3072} x layers {4,6,8} x heads {4,8,16}."""
    if not 1 <= conf <= 18:
        raise KeyError(conf)
    i = conf - 1
    return get_config(16, (2048, 3072)[i // 9], (4, 6, 8)[(i % 9) // 3], 256, (4, 8, 16)[i % 3], dropout_rate)

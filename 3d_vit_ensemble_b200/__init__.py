"""vit3d-b200: B200-native (sm_100a) forward/backward of the 3D-ViT stacking ensemble of
evapachetti/3d_vit_ensemble behind the reference's `models.modeling` class surface.

The directory name starts with a digit, so import it through the alias package::

    import vit3d_b200                      # = importlib.import_module("3d_vit_ensemble_b200")
    from vit3d_b200.models.modeling import VisionTransformer, TransformerEnsemble

or put this directory on sys.path and keep the reference's own import line
(`from models.modeling import VisionTransformer`).
"""
from . import _lib
from ._lib import Vit3dError, build
from .config import get_config, north_star_config, parameters_config
from .functional import get_precision, set_precision

__all__ = ["Vit3dError", "build", "get_config", "north_star_config", "parameters_config", "get_precision",
           "set_precision"]

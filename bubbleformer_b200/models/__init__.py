"""API mirror of upstream bubbleformer/models/__init__.py (the UNet baselines are outside the hot path)."""
from .axial_vit import *          # noqa: F401,F403
from .axial_vit import AViT, FiLMConditionedAViT, SpaceTimeBlock  # noqa: F401
from ._api import *               # noqa: F401,F403
from ._api import MODELS, get_model, list_models, register_model  # noqa: F401

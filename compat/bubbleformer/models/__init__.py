"""Mirror of upstream bubbleformer/models/__init__.py for the hot-path models."""
from bubbleformer_b200.models import *  # noqa: F401,F403
from bubbleformer_b200.models import MODELS, AViT, FiLMConditionedAViT, SpaceTimeBlock, get_model, list_models, register_model  # noqa: F401

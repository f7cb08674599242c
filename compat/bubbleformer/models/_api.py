"""Mirror of upstream bubbleformer/models/_api.py."""
from bubbleformer_b200.models._api import MODELS, get_model, list_models, register_model  # noqa: F401

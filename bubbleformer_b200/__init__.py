"""bubbleformer_b200: B200-native (sm_100a) implementation of the Bubbleformer FiLMAViT hot path.

Importing the package loads libbubbleformer_b200.so and raises if it has not been built
(`python bubbleformer_b200/build.py`); there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)
from .models import MODELS, get_model, list_models, register_model  # noqa: F401

__version__ = "0.1.0"

"""bubbleformer_b200: B200-native (sm_100a) implementation of the Bubbleformer FiLMAViT hot path.

Importing the package loads libbubbleformer_b200.so and raises if it has not been built
(`python bubbleformer_b200/build.py`); there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)
from .models import MODELS, get_model, list_models, register_model  # noqa: F401

__version__ = "0.1.0"


def set_exact_mode(on: bool = True) -> None:
    """Switch the whole library to the fp32 validation configuration (see engine.set_exact_mode); BF_EXACT=1 in the
    environment does the same at import."""
    from . import engine
    engine.set_exact_mode(on)


import os as _os

if _os.environ.get("BF_EXACT", "0") == "1":
    set_exact_mode(True)

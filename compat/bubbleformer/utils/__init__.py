"""`bubbleformer.utils` names that the B200 implementation provides (upstream bubbleformer/utils/__init__.py)."""
from .losses import LpLoss, eikonal_loss  # noqa: F401
from .heatflux import heatflux  # noqa: F401

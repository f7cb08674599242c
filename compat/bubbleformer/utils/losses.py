"""`bubbleformer.utils.losses` (upstream utils/losses.py) on the fused CUDA passes."""
from bubbleformer_b200.losses import LpLoss  # noqa: F401
from bubbleformer_b200.metrics import eikonal_loss  # noqa: F401

"""API mirror of upstream bubbleformer/layers/__init__.py (UNet conv blocks are outside the hot path)."""
from .positional_encoding import ContinuousPositionBias1D, RelativePositionBias
from .linear_layers import GeluMLP, SirenMLP, FiLMMLP
from .patching import HMLPEmbed, HMLPDebed
from .attention import AxialAttentionBlock, AttentionBlock

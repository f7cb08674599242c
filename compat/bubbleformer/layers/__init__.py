"""Mirror of upstream bubbleformer/layers/__init__.py (hot-path layers)."""
from bubbleformer_b200.layers import (AttentionBlock, AxialAttentionBlock, ContinuousPositionBias1D, FiLMMLP, GeluMLP,  # noqa: F401
                                      HMLPDebed, HMLPEmbed, RelativePositionBias, SirenMLP)

"""Mirror of upstream bubbleformer/models/axial_vit.py."""
from bubbleformer_b200.models.axial_vit import AViT, FiLMConditionedAViT, SpaceTimeBlock  # noqa: F401

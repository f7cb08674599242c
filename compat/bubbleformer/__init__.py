"""Drop-in `bubbleformer` package backed by bubbleformer_b200 (put `<repo>/compat` on sys.path / PYTHONPATH).

Upstream callers (`scripts/train.py:18`, `bubbleformer/modules.py:13`, `scripts/inference.py:4`) import
`bubbleformer.models.get_model`, `bubbleformer.models.axial_vit.SpaceTimeBlock` and the layer classes of
`bubbleformer.layers`; those names resolve here to the B200-native implementations, and so do the callers on either
side of the path: `bubbleformer.utils.losses` (LpLoss, eikonal_loss), `bubbleformer.utils.heatflux`, and
`bubbleformer.data.BubbleForecast` (HBM-resident trajectories).  Everything else (UNets, Lightning modules, plotting,
schedulers) stays upstream's and is not shadowed: keep upstream's own package for those.
"""
from bubbleformer_b200 import __version__  # noqa: F401

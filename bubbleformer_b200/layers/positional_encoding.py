"""Position-bias parameter holders (API mirror of upstream bubbleformer/layers/positional_encoding.py).

Inside the attention blocks the bias is never materialised: the fused attention kernel gathers
`relative_attention_bias.weight[bucket(j - i), head]` on the fly (bucket table from
`bubbleformer_b200.engine.relpos_bucket_vector`).  `forward` is kept for API compatibility.
"""
import torch
import torch.nn as nn

from ..engine import relpos_bucket_vector


class RelativePositionBias(nn.Module):
    """T5 relative position bias: 32 buckets x n_heads embedding (upstream positional_encoding.py:50-172)."""

    def __init__(self, bidirectional: bool = True, num_buckets: int = 32, max_distance: int = 128, n_heads: int = 2):
        super().__init__()
        if not bidirectional or num_buckets != 32:
            raise NotImplementedError("only the bidirectional 32-bucket table used by the AViT models is supported")
        self.bidirectional = bidirectional
        self.num_buckets = num_buckets
        self.max_distance = max_distance      # kept for parity; upstream never forwards it (static default 32 runs)
        self.n_heads = n_heads
        self.relative_attention_bias = nn.Embedding(self.num_buckets, self.n_heads)

    def forward(self, qlen: int, klen: int) -> torch.Tensor:
        """(1, n_heads, qlen, klen) bias tensor (compatibility path; not used by the fused kernels)."""
        w = self.relative_attention_bias.weight
        Lm = max(qlen, klen)
        vec = relpos_bucket_vector(Lm, w.device).long()           # index rel + Lm - 1
        i = torch.arange(qlen, device=w.device)[:, None]
        j = torch.arange(klen, device=w.device)[None, :]
        return w[vec[j - i + Lm - 1]].permute(2, 0, 1).unsqueeze(0)


class ContinuousPositionBias1D(nn.Module):
    """Parameter-compatible holder; unreachable from every shipped model config (bias_type is never passed)."""

    def __init__(self, n_heads: int):
        super().__init__()
        self.num_heads = n_heads
        self.cpb_mlp = nn.Sequential(nn.Linear(1, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, n_heads, bias=False))

    def forward(self, h: int, h2: int) -> torch.Tensor:
        raise NotImplementedError("bias_type='continuous' is outside the B200 hot path (upstream never selects it)")

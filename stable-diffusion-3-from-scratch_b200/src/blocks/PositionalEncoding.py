"""Timestep embedding with the reference's module API (reference: src/blocks/PositionalEncoding.py:8-30).

denom_i = 10000^(2i/dim) for i in [0, dim); out = cat(sin(t/denom[0::2]), cos(t/denom[1::2])).
`denom` stays a plain attribute (not a buffer), as in the reference."""
import torch
from torch import nn

from mmdit.functional import TimestepEmbedFn


class PositionalEncoding(nn.Module):
    def __init__(self, dim, device):
        super().__init__()
        self.dim = dim
        self.denom = (torch.tensor(10000.0) ** ((2 * torch.arange(self.dim)) / self.dim)).to(
            dtype=torch.float, device=device)

    def embed(self, time, time_scale):
        """sin/cos embedding of time * time_scale (the multiply is fused into the kernel)."""
        if self.denom.device != time.device:
            self.denom = self.denom.to(time.device)
        return TimestepEmbedFn.apply(time, time_scale, self.denom)

    def forward(self, time):
        one = torch.ones(1, device=time.device, dtype=torch.float32)
        return self.embed(time, one)

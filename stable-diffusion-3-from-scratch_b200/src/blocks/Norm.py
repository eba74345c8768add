"""adaLN normalisation with the reference's module API (reference: src/blocks/Norm.py:5-23).

LayerNorm without affine parameters followed by X*(1+c_scale(y)) + c_shift(y).
Here the two modulation projections run as one packed GEMM and the
normalise + modulate step is one fused 128-bit-vectorised kernel.
"""
import torch
from torch import nn

from mmdit.functional import LinearFn, LNModulateFn, LNModulateResFn
from mmdit.shadow import packed_weight

BF16 = torch.bfloat16


def modulate(X, shift, scale):
    """LN(X) * (1 + scale[:, None]) + shift[:, None]; X [B,T,d], shift/scale [B,d] bf16 views."""
    return LNModulateFn.apply(X if X.dtype == BF16 else X.to(BF16), shift, scale)


def modulate_keep(X, shift, scale):
    """Like modulate, but also returns X for use as the residual (fuses the gradient sum)."""
    return LNModulateResFn.apply(X if X.dtype == BF16 else X.to(BF16), shift, scale)


class Norm(nn.Module):
    def __init__(self, dim, c_dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim, elementwise_affine=False)  # parameter-free; kept for API parity
        self.c_shift = nn.Linear(c_dim, dim, bias=False)
        self.c_scale = nn.Linear(c_dim, dim, bias=False)

    def weights(self):
        return [self.c_shift.weight, self.c_scale.weight]

    def forward(self, X, y=None):
        d = self.c_shift.weight.shape[0]
        wb = packed_weight(self, "mod", self.weights())
        mod = LinearFn.apply(y.to(BF16), wb, None, 0, 2, *self.weights())  # [B, 2d]
        return modulate(X, mod[:, :d], mod[:, d:])

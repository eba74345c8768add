"""SwiGLU feed-forward with the reference's module API (reference: src/blocks/MLP.py:7-40,
which wraps xformers.ops.swiglu_op.SwiGLU 0.0.29.post3: packed w12 = [w1; w2], bias on both
linears, out = w3(silu(x1) * x2) with x1, x2 = w12(x).chunk(2)).  state_dict keys:
MLP.w12.{weight,bias}, MLP.w3.{weight,bias}."""
import torch
from torch import nn

from mmdit.functional import LinearFn, SwiGLUHiddenFn
from mmdit.shadow import packed_weight

BF16 = torch.bfloat16


class SwiGLU(nn.Module):
    def __init__(self, in_features, hidden_features, out_features=None, bias=True):
        super().__init__()
        out_features = out_features or in_features
        self.w12 = nn.Linear(in_features, 2 * hidden_features, bias=bias)
        self.w3 = nn.Linear(hidden_features, out_features, bias=bias)
        self.hidden_features = hidden_features

    def hidden(self, X):
        """silu(x1) * x2 -- everything before w3 (the block fuses w3 with gate + residual)."""
        wb = packed_weight(self, "w12", [self.w12.weight])
        return SwiGLUHiddenFn.apply(X if X.dtype == BF16 else X.to(BF16), wb, self.w12.weight, self.w12.bias)

    def forward(self, X):
        a = self.hidden(X)
        wb = packed_weight(self, "w3", [self.w3.weight])
        params = [self.w3.weight] + ([self.w3.bias] if self.w3.bias is not None else [])
        return LinearFn.apply(a, wb, None if self.w3.bias is None else self.w3.bias.detach(), 0, 1,
                              *params)


class MLP(nn.Module):
    def __init__(self, dim, hidden_scale=4.0, act="swiglu"):
        super().__init__()
        self.proj_size = int(dim * hidden_scale)
        self.act_ = act
        if act != "swiglu":
            raise NotImplementedError(
                f"MLP act={act!r}: only the hot-path 'swiglu' MLP is built on the B200 path "
                "(reference train.py:41); 'gelu' is out of scope (SURVEY 2.1)")
        self.MLP = SwiGLU(dim, self.proj_size, dim)

    def forward(self, X):
        return self.MLP(X)

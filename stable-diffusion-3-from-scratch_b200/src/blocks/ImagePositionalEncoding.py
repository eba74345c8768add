"""Patch embedding with the reference's module API (reference:
src/blocks/ImagePositionalEncoding.py:90-203).  The stride-p conv is a GEMM over
patch vectors: tokens [B*N, C*p*p] x proj.weight.view(dim, C*p*p)^T.  Only the
position-embedding-free modes used with rotary attention are built."""
import torch
from torch import nn

from mmdit.functional import LinearFn, PatchifyFn
from mmdit.shadow import packed_weight


class PatchEmbed(nn.Module):
    def __init__(self, height=224, width=224, patch_size=16, in_channels=3, embed_dim=768,
                 layer_norm=False, flatten=True, bias=True, interpolation_scale=1,
                 pos_embed_type="absolute", pos_embed_max_size=None):
        super().__init__()
        if pos_embed_type not in (None, "RoPE", "NoPE", "RoPE2d", "RoPE2dV2"):
            raise NotImplementedError(
                f"PatchEmbed pos_embed_type={pos_embed_type!r}: the sincos 'absolute' table is not on the "
                "hot path (reference train.py:63 uses RoPE2d)")
        if layer_norm or bias or not flatten:
            raise NotImplementedError("PatchEmbed: layer_norm / bias / flatten=False are not used by diff_model")
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size),
                              stride=patch_size, bias=False)
        self.patch_size = patch_size
        self.height, self.width = height // patch_size, width // patch_size
        self.pos_embed = None
        self.pos_embed_max_size = pos_embed_max_size

    def forward(self, latent):
        B, C, H, W = latent.shape
        p = self.patch_size
        tok = PatchifyFn.apply(latent, p)                      # [B*N, C*p*p] bf16
        wb = packed_weight(self, "proj", [self.proj.weight])   # [dim, C*p*p]
        x = LinearFn.apply(tok, wb, None, 0, 1, self.proj.weight)
        return x.view(B, (H // p) * (W // p), -1)

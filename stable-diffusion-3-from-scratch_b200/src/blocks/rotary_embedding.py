"""Axial rotary embedding tables (reference: src/blocks/rotary_embedding.py:90-166 ctor,
:269-288 get_axial_freqs, :290-321 forward, :36-76 rotate_half/apply_rotary_emb).

Only what the MMDiT hot path uses is kept: the frozen `freqs` parameter (it is part of the
state_dict), the (h, w, head_dim) angle grid, and cos/sin tables per
interleaved pair that the fused QK-norm+RoPE kernel reads."""
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000, interpolate_factor=1.0, use_xpos=False, **_unused):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)
        # accepted and stored like the reference; get_axial_freqs ignores it there too
        self.interpolate_factor = interpolate_factor
        self._tables = {}

    def get_axial_freqs(self, *dims):
        """(d1, d2, ..., 2*len(freqs)*len(dims)) angles; axis k uses positions arange(dims[k])."""
        parts = []
        for ind, n in enumerate(dims):
            pos = torch.arange(n, device=self.freqs.device)
            f = torch.einsum("i,f->if", pos.type(self.freqs.dtype), self.freqs)
            f = f.repeat_interleave(2, dim=-1)
            shape = [1] * len(dims) + [f.shape[-1]]
            shape[ind] = n
            parts.append(f.view(shape).expand(*dims, f.shape[-1]))
        return torch.cat(parts, dim=-1)

    def tables(self, h, w):
        """cos/sin of the angle of every interleaved pair: two fp32 [h*w, 32] tensors."""
        key = (h, w, self.freqs.device, self.freqs._version, self.freqs.data_ptr())
        t = self._tables.get(key)
        if t is None:
            with torch.no_grad():
                ang = self.get_axial_freqs(h, w)[..., 0::2].reshape(h * w, -1).float()
                t = (ang.cos().contiguous(), ang.sin().contiguous())
            self._tables = {key: t}
        return t


def rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def apply_rotary_emb(freqs, t, start_index=0, scale=1.0, seq_dim=-2, freqs_seq_dim=None):
    """Eager restatement kept for API parity (used by tests, not by the hot path)."""
    dtype = t.dtype
    rot_dim = freqs.shape[-1]
    mid = t[..., start_index:start_index + rot_dim]
    out = (mid * freqs.cos() * scale) + (rotate_half(mid) * freqs.sin() * scale)
    return torch.cat((t[..., :start_index], out, t[..., start_index + rot_dim:]), dim=-1).type(dtype)

"""Joint text+image attention with the reference's module API (reference:
src/blocks/Attention.py:15-113 ctor, :118-427 forward).

Built path (the one BASELINE.json names): dual streams, softmax attention with per-head
QK RMSNorm, 2-D axial RoPE on image tokens, image-then-text joint sequence, head_dim 64.
  * query|key|value of a stream run as ONE packed GEMM ([3d, d] bf16 shadow weight),
  * QK-RMSNorm + RoPE is one fused kernel on the packed projections,
  * the attention kernel reads Q/K/V of both streams in place (no concat / transpose /
    split copies) and writes per-stream [rows, d] outputs,
  * the out-projections are GEMMs (the block fuses them with gate + residual).
Other attn_type / positional_encoding / kv_merge_attn / qk_half_dim variants of the
reference are research flags outside the hot path and raise NotImplementedError.
"""
import torch
from torch import nn

from mmdit.functional import JointAttentionFn, LinearFn
from mmdit.shadow import packed_weight
from src.blocks.rotary_embedding import RotaryEmbedding

BF16 = torch.bfloat16


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, attn_type="cosine", causal=False, emb_dim=None,
                 positional_encoding="absolute", RoPE_Scale=1, kv_merge_attn=False, qk_half_dim=False,
                 layer_idx=None, dual=False, last=False):
        super().__init__()
        if attn_type not in ("softmax", "softmax_flash"):
            raise NotImplementedError(f"attn_type={attn_type!r}: only softmax / softmax_flash are on the B200 path")
        if not dual or causal or emb_dim is not None or kv_merge_attn or qk_half_dim:
            raise NotImplementedError("Attention: only the dual-stream, non-causal, full-width configuration is built")
        if positional_encoding not in ("RoPE2d", "NoPE"):
            raise NotImplementedError(f"positional_encoding={positional_encoding!r}: RoPE2d (or NoPE) only")
        if dim % num_heads or dim // num_heads != 64:
            raise NotImplementedError("Attention kernels are built for head_dim == 64 (reference train.py:36-39)")
        self.positional_encoding = positional_encoding
        self.kv_merge_attn = kv_merge_attn
        self.RoPE_Scale = RoPE_Scale
        self.layer_idx = layer_idx
        self.dual = dual
        self.last = last
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = self.head_dim_qk = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.attn_type = attn_type
        self.causal = causal

        self.query_proj_x = nn.Linear(dim, dim, bias=False)
        self.key_proj_x = nn.Linear(dim, dim, bias=False)
        self.value_proj_x = nn.Linear(dim, dim, bias=False)
        self.out_proj_x = nn.Linear(dim, dim, bias=False)
        self.query_proj_c = nn.Linear(dim, dim, bias=False)
        self.key_proj_c = nn.Linear(dim, dim, bias=False)
        self.value_proj_c = nn.Linear(dim, dim, bias=False)
        if not self.last:
            self.out_proj_c = nn.Linear(dim, dim, bias=False)
        self.q_norm_x = nn.RMSNorm(self.head_dim_qk)
        self.k_norm_x = nn.RMSNorm(self.head_dim_qk)
        self.q_norm_c = nn.RMSNorm(self.head_dim_qk)
        self.k_norm_c = nn.RMSNorm(self.head_dim_qk)
        if positional_encoding == "RoPE2d":
            # half of the head per axis (Attention.py:96-98)
            self.rotary_emb = RotaryEmbedding(self.head_dim_qk // 2, use_xpos=False,
                                              interpolate_factor=1 / RoPE_Scale)

    def attend(self, x, c, orig_shape):
        """Everything up to (not including) the output projections.
        x [B,N,d], c [B,M,d] -> (a_x [B*N,d], a_c [B*M,d]) bf16."""
        B, N, d = x.shape
        M = c.shape[1]
        return self.attend_qkv(self.project_qkv(x, "x"), self.project_qkv(c, "c"), orig_shape, B, N, M)

    def project_qkv(self, t, which):
        """Packed q|k|v projection of one stream ("x": image tokens, "c": text tokens) -> [B*T, 3d]."""
        B, T, d = t.shape
        t = t if t.dtype == BF16 else t.to(BF16)
        if which == "x":
            ws = [self.query_proj_x.weight, self.key_proj_x.weight, self.value_proj_x.weight]
        else:
            ws = [self.query_proj_c.weight, self.key_proj_c.weight, self.value_proj_c.weight]
        return LinearFn.apply(t.reshape(B * T, d), packed_weight(self, "qkv_" + which, ws), None, 0, 3, *ws)

    def project_qkv_prenorm(self, t, which, orig_shape):
        """Experimental (functional.FUSED_QKNORM): the projection GEMM's epilogue also produces the
        normalised / rotated q, k.  Returns (qkv [B*T,3d], qk [B*T,2d])."""
        from mmdit.functional import QKVProjFn
        B, T, d = t.shape
        t = t if t.dtype == BF16 else t.to(BF16)
        cos = sin = None
        if which == "x":
            ws = [self.query_proj_x.weight, self.key_proj_x.weight, self.value_proj_x.weight]
            wq, wk = self.q_norm_x.weight, self.k_norm_x.weight
            if self.positional_encoding == "RoPE2d":
                cos, sin = self.rotary_emb.tables(orig_shape[-2] // 2, orig_shape[-1] // 2)
        else:
            ws = [self.query_proj_c.weight, self.key_proj_c.weight, self.value_proj_c.weight]
            wq, wk = self.q_norm_c.weight, self.k_norm_c.weight
        return QKVProjFn.apply(t.reshape(B * T, d), packed_weight(self, "qkv_" + which, ws), wq, wk,
                               cos, sin, T, *ws)

    def attend_prenorm(self, x_pair, c_pair, orig_shape, B, N, M):
        """Joint attention on (qkv, qk) pairs from project_qkv_prenorm."""
        from mmdit.functional import JointAttentionPreNormFn
        cos = sin = None
        if self.positional_encoding == "RoPE2d":
            cos, sin = self.rotary_emb.tables(orig_shape[-2] // 2, orig_shape[-1] // 2)
        return JointAttentionPreNormFn.apply(x_pair[0], c_pair[0], x_pair[1], c_pair[1],
                                             self.q_norm_x.weight, self.k_norm_x.weight,
                                             self.q_norm_c.weight, self.k_norm_c.weight, cos, sin,
                                             B, self.num_heads, N, M)

    def attend_qkv(self, qkv_x, qkv_c, orig_shape, B, N, M):
        """QK-norm + RoPE + joint attention on the packed projections of both streams."""
        cos = sin = None
        if self.positional_encoding == "RoPE2d":
            # patch size 2 is hard-coded in the reference as well (Attention.py:178-179)
            h, w = orig_shape[-2] // 2, orig_shape[-1] // 2
            if h * w != N:
                raise ValueError(f"RoPE2d: {N} image tokens do not form a {h}x{w} grid")
            cos, sin = self.rotary_emb.tables(h, w)
        return JointAttentionFn.apply(qkv_x, qkv_c, self.q_norm_x.weight, self.k_norm_x.weight,
                                      self.q_norm_c.weight, self.k_norm_c.weight, cos, sin,
                                      B, self.num_heads, N, M)

    def forward(self, x, c=None, orig_shape=None):
        assert c is not None, "Dual attention requires context tensor c"
        B, N, d = x.shape
        M = c.shape[1]
        a_x, a_c = self.attend(x, c, orig_shape)
        wo = self.out_proj_x.weight
        out_x = LinearFn.apply(a_x, packed_weight(self, "out_x", [wo]), None, 0, 1, wo).view(B, N, d)
        if self.last:
            return out_x, a_c.view(B, M, d)
        wo = self.out_proj_c.weight
        out_c = LinearFn.apply(a_c, packed_weight(self, "out_c", [wo]), None, 0, 1, wo).view(B, M, d)
        return out_x, out_c

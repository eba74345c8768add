"""One MMDiT dual-stream block with the reference's module API (reference:
src/blocks/Transformer_Block_Dual.py:15-53 ctor, :56-78 forward).

Same parameters and math as the reference block; the schedule is different:
  * the 13 adaLN projections of y' (8 shift/scale, 4 gates; 9 in the last block) run as
    ONE GEMM against a packed [13d, d] shadow weight,
  * LN + modulate is one kernel per use,
  * each out-projection / SwiGLU-w3 GEMM applies `* gate + residual` in its epilogue.
checkpoint_MLP / checkpoint_attn are accepted for API compatibility; activations are kept
instead of recomputed (180 GB of HBM3e makes recompute a pure loss at these sizes).
"""
import torch
from torch import nn

from mmdit import ops, streams
from mmdit import functional as Fn
from mmdit.functional import FUSED_QKNORM, GatedLinearFn, GatedLinearLNFn, LinearFn
from mmdit.shadow import packed_weight
from src.blocks.Attention import Attention
from src.blocks.MLP import MLP, SwiGLU
from src.blocks.Norm import Norm, modulate, modulate_keep

BF16 = torch.bfloat16


class Transformer_Block_Dual(nn.Module):
    def __init__(self, dim, c_dim, hidden_scale=4.0, num_heads=8, attn_type="softmax", MLP_type="gelu",
                 causal=False, positional_encoding="absolute", RoPE_Scale=1, kv_merge_attn=False,
                 qk_half_dim=False, checkpoint_MLP=True, checkpoint_attn=True, layer_idx=None, last=False):
        super().__init__()
        if c_dim != dim:
            raise NotImplementedError("Transformer_Block_Dual: c_dim == dim (as diff_model builds it)")
        self.checkpoint_MLP = checkpoint_MLP
        self.checkpoint_attn = checkpoint_attn
        self.last = last
        self.dim = dim
        self.y_proj = nn.Sequential(nn.Linear(c_dim, c_dim), nn.SiLU())
        if MLP_type == "swiglu_old":
            self.MLP_x = SwiGLU(dim, int(dim * hidden_scale), dim)
            if not self.last:
                self.MLP_c = SwiGLU(dim, int(dim * hidden_scale), dim)
        else:
            self.MLP_x = MLP(dim, hidden_scale, act=MLP_type)
            if not self.last:
                self.MLP_c = MLP(dim, hidden_scale, act=MLP_type)
        self.attn = Attention(dim, num_heads=num_heads, attn_type=attn_type, causal=causal,
                              positional_encoding=positional_encoding, RoPE_Scale=RoPE_Scale,
                              kv_merge_attn=kv_merge_attn, qk_half_dim=qk_half_dim, layer_idx=layer_idx,
                              dual=True, last=last)
        self.norm1_x = Norm(dim, c_dim)
        self.norm2_x = Norm(dim, c_dim)
        self.norm1_c = Norm(dim, c_dim)
        if not self.last:
            self.norm2_c = Norm(dim, c_dim)
        self.scale1_x = nn.Linear(c_dim, dim, bias=False)
        self.scale2_x = nn.Linear(c_dim, dim, bias=False)
        if not self.last:
            self.scale1_c = nn.Linear(c_dim, dim, bias=False)
            self.scale2_c = nn.Linear(c_dim, dim, bias=False)

    # order of the packed modulation GEMM's output slots
    def _mod_weights(self):
        ws = [self.norm1_x.c_shift.weight, self.norm1_x.c_scale.weight,      # 0 1
              self.norm1_c.c_shift.weight, self.norm1_c.c_scale.weight,      # 2 3
              self.scale1_x.weight,                                          # 4
              self.norm2_x.c_shift.weight, self.norm2_x.c_scale.weight,      # 5 6
              self.scale2_x.weight]                                          # 7
        if not self.last:
            ws += [self.scale1_c.weight,                                     # 8
                   self.norm2_c.c_shift.weight, self.norm2_c.c_scale.weight, # 9 10
                   self.scale2_c.weight]                                     # 11
        return ws

    @staticmethod
    def _swiglu(mlp):
        return mlp.MLP if isinstance(mlp, MLP) else mlp

    def _gated(self, a, lin, gate, resid, rows_per_batch):
        """resid + gate * lin(a) in one GEMM."""
        wb = packed_weight(lin, "w", [lin.weight])
        bb = None if lin.bias is None else lin.bias.detach()
        B, T, d = resid.shape
        o = GatedLinearFn.apply(a, wb, bb, gate, resid.reshape(B * T, d), rows_per_batch,
                                lin.weight, lin.bias)
        return o.view(B, T, d)

    def forward(self, X, c, y, orig_shape, yp=None):
        """yp (optional): SiLU(y_proj(y)) computed by the caller for all blocks at once
        (diff_model batches the per-block y projections into one GEMM); by default it is
        computed here, as in the reference."""
        B, N, d = X.shape
        M = c.shape[1]
        X = X if X.dtype == BF16 else X.to(BF16)
        c = c if c.dtype == BF16 else c.to(BF16)
        if yp is None:
            lin = self.y_proj[0]
            yp = LinearFn.apply(y if y.dtype == BF16 else y.to(BF16), packed_weight(lin, "w", [lin.weight]),
                                lin.bias.detach(), ops.EPI_SILU, 1, lin.weight, lin.bias)
        ws = self._mod_weights()
        mod = LinearFn.apply(yp, packed_weight(self, "mod", ws), None, 0, len(ws), *ws)
        m = mod.unflatten(1, (len(ws), d)).unbind(1)

        # modulate_keep returns (LN-mod(X), X): taking the residual from the second output lets the LN
        # backward kernel add the residual-path gradient itself (no separate elementwise add)
        if streams.active(X):
            return self._forward_two_streams(X, c, m, orig_shape, B, N, M)
        xn, X = modulate_keep(X, m[0], m[1])
        if self.last:
            cn = modulate(c, m[2], m[3])
        else:
            cn, c = modulate_keep(c, m[2], m[3])
        a_x, a_c = self.attn.attend(xn, cn, orig_shape)
        X = self._x_branch(a_x, X, m, B, N)
        if not self.last:
            c = self._c_branch(a_c, c, m, B, M)
        return X, c

    def _gated_ln(self, a, lin, gate, resid, shift, scale, rows_per_batch):
        """(LN-mod(X'), X') with X' = resid + gate * lin(a): the out-projection's gated residual and the
        LayerNorm-modulate in front of the MLP share one pass over the GEMM output."""
        # wider rows than the one-pass kernel is built for go through the two kernels (first generation:
        # 1024 columns, second: 1536; ops.fused_gate_ln_max_columns)
        if not Fn.FUSED_GATE_LN or resid.shape[-1] > ops.fused_gate_ln_max_columns():
            return modulate_keep(self._gated(a, lin, gate, resid, rows_per_batch), shift, scale)
        wb = packed_weight(lin, "w", [lin.weight])
        bb = None if lin.bias is None else lin.bias.detach()
        B, T, d = resid.shape
        y, xo = GatedLinearLNFn.apply(a, wb, bb, gate, resid.reshape(B * T, d), shift, scale, rows_per_batch,
                                      lin.weight, lin.bias)
        return y.view(B, T, d), xo.view(B, T, d)

    def _x_branch(self, a_x, X, m, B, N):
        """Image stream after the attention: out-projection, gate, MLP, gate."""
        xn, X = self._gated_ln(a_x, self.attn.out_proj_x, m[4], X, m[5], m[6], N)
        mx = self._swiglu(self.MLP_x)
        return self._gated(mx.hidden(xn).reshape(B * N, -1), mx.w3, m[7], X, N)

    def _c_branch(self, a_c, c, m, B, M):
        """Text stream after the attention (absent in the last block)."""
        cn, c = self._gated_ln(a_c, self.attn.out_proj_c, m[8], c, m[9], m[10], M)
        mc = self._swiglu(self.MLP_c)
        return self._gated(mc.hidden(cn).reshape(B * M, -1), mc.w3, m[11], c, M)

    def _forward_two_streams(self, X, c, m, orig_shape, B, N, M):
        """Same math, text branch on the side stream (mmdit/streams.py): the branches only meet at
        the joint attention, so each block is fork -> [LN+QKV] x2 -> join -> attention -> fork."""
        main = torch.cuda.current_stream()
        side = streams.side(X.device)
        side.wait_stream(main)                      # modulation (and the incoming c) are ready
        with torch.cuda.stream(side):
            if self.last:
                cn = modulate(c, m[2], m[3])
            else:
                cn, c = modulate_keep(c, m[2], m[3])
            fused_qk = FUSED_QKNORM and B * M > 128 and B * N > 128     # experimental, off by default
            qkv_c = (self.attn.project_qkv_prenorm(cn, "c", orig_shape) if fused_qk
                     else self.attn.project_qkv(cn, "c"))
        xn, X = modulate_keep(X, m[0], m[1])
        if fused_qk:
            qkv_x = self.attn.project_qkv_prenorm(xn, "x", orig_shape)
            main.wait_stream(side)
            a_x, a_c = self.attn.attend_prenorm(qkv_x, qkv_c, orig_shape, B, N, M)
            streams.hold(*m, *qkv_c, a_c, c)
        else:
            qkv_x = self.attn.project_qkv(xn, "x")
            main.wait_stream(side)                      # text q|k|v ready
            a_x, a_c = self.attn.attend_qkv(qkv_x, qkv_c, orig_shape, B, N, M)
            streams.hold(*m, qkv_c, a_c, c)             # no_grad only: tensors both streams touch
        if not self.last:
            side.wait_stream(main)                  # attention output ready
            with torch.cuda.stream(side):
                c = self._c_branch(a_c, c, m, B, M)
                streams.hold(c)
        X = self._x_branch(a_x, X, m, B, N)
        if self.last:
            main.wait_stream(side)                  # nothing of the text branch is left in flight
        return X, c

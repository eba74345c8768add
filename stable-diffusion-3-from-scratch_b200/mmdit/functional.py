"""torch.autograd.Function glue between the reference-shaped modules (src/) and
the C-ABI kernels (ops.py).  Parameters stay fp32 nn.Parameters (state_dict
contract); the kernels consume bf16 "shadow" copies that the
modules provide, and return fp32 weight gradients through autograd, so
torch.optim / GradScaler / clip_grad_norm_ / DDP keep working unchanged.

Backward math follows SURVEY App. E (restated from the reference forward; the
reference itself relies on autograd).
"""
import os

import torch
from torch.autograd import Function

from . import ops, streams, zeropool
from .ops import BF16, F32


# MMDIT_FUSED_GATE=1 routes gate*x+residual through the GEMM epilogue instead of a separate kernel
FUSED_GATE_EPILOGUE = os.environ.get("MMDIT_FUSED_GATE", "0") == "1"
# MMDIT_FUSED_GATE_LN=0 keeps the gated residual and the LayerNorm-modulate that follows it in two kernels
FUSED_GATE_LN = os.environ.get("MMDIT_FUSED_GATE_LN", "1") == "1"
# MMDIT_FUSED_SWIGLU=0 keeps silu(x1)*x2 in its own kernel instead of the w12 GEMM's epilogue
FUSED_SWIGLU = os.environ.get("MMDIT_FUSED_SWIGLU", "1") == "1"
# MMDIT_FUSED_QKNORM=1 (experimental, not yet validated on hardware): per-head RMSNorm + 2-D RoPE of
# q and k in the epilogue of the packed q|k|v projection (QKVProjFn / JointAttentionPreNormFn below)
FUSED_QKNORM = os.environ.get("MMDIT_FUSED_QKNORM", "0") == "1"
# MMDIT_FUSED_SWIGLU_BWD=1: the SwiGLU backward runs in the epilogue of w3's data-gradient GEMM
FUSED_SWIGLU_BWD = os.environ.get("MMDIT_FUSED_SWIGLU_BWD", "0") == "1"


class _SwigluLink:
    """Shared by a SwiGLUHiddenFn node and the gated-linear node (w3) that consumes its output, so that
    the w3 data-gradient GEMM can apply the SwiGLU backward in its epilogue (ops.gemm_swiglu_bwd).
    The w3 node hands dh12 back through this object; the gradient it returns for the activation is a
    correctly shaped view of dh12 that SwiGLUHiddenFn.backward recognises by its address and ignores."""
    __slots__ = ("h12", "has_bias", "paired", "dh12", "db", "dummy_ptr")

    def __init__(self, h12, has_bias):
        self.h12, self.has_bias, self.paired = h12, has_bias, False
        self.dh12 = self.db = self.dummy_ptr = None


_SWIGLU_LINKS = {}   # data_ptr of a SwiGLU activation -> its link, until the consumer's forward claims it


def _claim_swiglu_link(a):
    link = _SWIGLU_LINKS.pop(a.data_ptr(), None) if _SWIGLU_LINKS else None
    if link is not None:
        link.paired = True
    return link


def _w3_dgrad(link, da, wb):
    """dL/d(input of w3): plain GEMM, or -- when the input is a SwiGLU activation -- the fused GEMM that
    already yields dh12 (returned to the SwiGLU node through the link; the caller gets a dummy view)."""
    if link is None:
        return ops.gemm(da, wb, b_major=1)
    hid = wb.shape[1]
    link.db = zeropool.zeros(2 * hid, da.device) if link.has_bias else None
    link.dh12 = ops.gemm_swiglu_bwd(da, wb, link.h12, link.db)
    dummy = link.dh12[:, :hid]
    link.dummy_ptr = dummy.data_ptr()
    return dummy


def _wgrad_slot(params):
    """Data parallelism (mmdit.train.GradBuckets) gives every parameter a slot inside a flat fp32
    gradient bucket.  When the parameters of one packed GEMM sit back to back in that bucket, the
    wgrad GEMM writes the bucket directly: returns the fp32 [rows, K] matrix aliasing the slots
    (None otherwise).  autograd then adopts the returned views as .grad -- no zero-fill of the
    bucket, no accumulate pass, no copy."""
    slots = [getattr(p, "_grad_slot", None) for p in params]
    if not slots or any(s is None for s in slots):
        return None
    ptr = slots[0].data_ptr()
    if ptr % 16:
        return None
    for s_ in slots:
        if s_.data_ptr() != ptr:
            return None
        ptr += s_.numel() * 4
    rows = sum(p.shape[0] for p in params)
    K = params[0][0].numel()
    return torch.as_strided(slots[0], (rows, K), (K, 1), slots[0].storage_offset())


def _wgrad_gemm(dy2, x2, params):
    """dW = dY^T X (fp32), written into the parameters' gradient-bucket slot when they have one.
    Under the trainer (streams.async_wgrad) the GEMM runs on the weight-gradient stream: nothing
    in the backward depends on it, so the data-gradient chain goes on without waiting."""
    out = _wgrad_slot(params)
    if not (streams.async_wgrad and dy2.is_cuda):
        return ops.gemm(dy2, x2, a_major=1, b_major=1, out_dtype=F32, out=out)
    ws = streams.wgrad(dy2.device)
    ws.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(ws):
        dw = ops.gemm(dy2, x2, a_major=1, b_major=1, out_dtype=F32, out=out)
    streams.keepalive.append((dy2, x2))
    return dw


def _split_rows(t, sizes):
    """Row slices of a packed [sum(sizes), K] tensor (views, no copies)."""
    out, r = [], 0
    for n in sizes:
        out.append(t[r:r + n])
        r += n
    return out


class LinearFn(Function):
    """y = act(x @ W^T + b); W is the row-concatenation of `nw` fp32 parameters
    (e.g. query/key/value projections, Attention.py:130-135), `wb`/`bb` are the
    bf16 weight / fp32 bias shadows.  act: 0 none, 2 SiLU (Transformer_Block_Dual.py:25-28)."""

    @staticmethod
    def forward(ctx, x, wb, bb, act, nw, *params):
        x2 = x.reshape(-1, x.shape[-1])
        aux = None
        if act == ops.EPI_SILU:  # pre-activation, needed by backward
            aux = torch.empty((x2.shape[0], wb.shape[0]), device=x.device, dtype=BF16)
        y = ops.gemm(x2, wb, bias=bb, epilogue=act, aux=aux)
        ctx.save_for_backward(x2, wb, aux)
        ctx.meta = (x.shape, act, nw, [p.shape[0] for p in params[:nw]], len(params) > nw)
        ctx.wshapes = [p.shape for p in params[:nw]]
        ctx.wparams = params[:nw]
        return y.reshape(*x.shape[:-1], wb.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb, aux = ctx.saved_tensors
        xshape, act, nw, sizes, has_bias = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1])
        if act == ops.EPI_SILU:
            z = aux.float()
            s = torch.sigmoid(z)
            dy2 = (dy2.float() * s * (1 + z * (1 - s))).to(BF16)
        elif not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy2, wb, b_major=1).reshape(xshape)
        dw = _wgrad_gemm(dy2, x2, ctx.wparams)
        grads = [g.view(shp) for g, shp in zip(_split_rows(dw, sizes), ctx.wshapes)]
        if has_bias:
            grads += list(_split_rows(ops.colsum(dy2), sizes))
        return (dx, None, None, None, None, *grads)


class GatedLinearFn(Function):
    """o = (a @ W^T + b) * gate[sample] + resid, one GEMM with a fused epilogue
    (out-projection / SwiGLU w3 followed by the adaLN-Zero gate and the residual,
    Transformer_Block_Dual.py:64-76).  a: [R,K] bf16, gate: [B,N] bf16 view,
    resid: [R,N] bf16.  The pre-gate value is kept (bf16) for dgate."""

    @staticmethod
    def forward(ctx, a, wb, bb, gate, resid, rows_per_batch, w, b):
        R = a.shape[0]
        if FUSED_GATE_EPILOGUE:
            aux = torch.empty((R, wb.shape[0]), device=a.device, dtype=BF16)
            o = ops.gemm(a, wb, bias=bb, epilogue=ops.EPI_GATE_RESID, gate=gate,
                         rows_per_gate=rows_per_batch, resid=resid, aux=aux)
        else:
            # measured on B200 (profiles/): the plain-epilogue GEMM plus one streaming pass beats the
            # fused gate+residual epilogue, whose dependent global loads stall the 4 epilogue warps
            aux = ops.gemm(a, wb, bias=bb)
            o = ops.gate_residual_fwd(aux, gate, resid, rows_per_batch)
        ctx.save_for_backward(a, wb, aux, gate)
        ctx.link = _claim_swiglu_link(a)
        ctx.wparam = w
        ctx.rpb = rows_per_batch
        ctx.has_bias = b is not None
        return o

    @staticmethod
    def backward(ctx, do):
        a, wb, aux, gate = ctx.saved_tensors
        rpb = ctx.rpb
        do = do.contiguous()
        Bn, n = gate.shape
        dgate = torch.empty((Bn, n), device=do.device, dtype=BF16)       # written by the kernel
        dab = torch.empty((Bn, n), device=do.device, dtype=F32) if ctx.has_bias else None
        da = ops.gate_bwd(do, aux, gate, dgate, dab, rpb)
        dx = _w3_dgrad(ctx.link, da, wb)
        dw = _wgrad_gemm(da, a, [ctx.wparam]).view(ctx.wparam.shape)
        db = dab.sum(0) if ctx.has_bias else None
        return dx, None, None, dgate, do, None, dw, db


class GatedLinearLNFn(Function):
    """GatedLinearFn followed by the next adaLN LayerNorm-modulate (Transformer_Block_Dual.py:64-72),
    with the gate, the residual add, the LayerNorm statistics and the modulate in ONE pass over the
    GEMM output: returns (LN-mod(X'), X') with X' = (a @ W^T + b) * gate + resid.
    The backward is LNModulateResFn's followed by GatedLinearFn's."""

    @staticmethod
    def forward(ctx, a, wb, bb, gate, resid, shift, scale, rows_per_batch, w, b):
        aux = ops.gemm(a, wb, bias=bb)
        xo, y, mean, rstd = ops.gate_residual_ln_fwd(aux, gate, resid, shift, scale, rows_per_batch)
        ctx.save_for_backward(a, wb, aux, gate, xo, mean, rstd, scale)
        ctx.link = _claim_swiglu_link(a)
        ctx.wparam = w
        ctx.rpb = rows_per_batch
        ctx.has_bias = b is not None
        ctx.set_materialize_grads(False)
        return y, xo

    @staticmethod
    def backward(ctx, dy, dres):
        a, wb, aux, gate, xo, mean, rstd, scale = ctx.saved_tensors
        rpb = ctx.rpb
        Bn, n = gate.shape
        dmod = None
        if dy is None:          # only the residual path was used
            do = dres.contiguous()
        else:
            dmod = torch.empty((2, Bn, n), device=dy.device, dtype=BF16)
            dr = None if dres is None else dres.contiguous()
            do = ops.ln_modulate_bwd(dy.contiguous(), xo, mean, rstd, scale, dr, dmod[0], dmod[1], rpb)
        dgate = torch.empty((Bn, n), device=do.device, dtype=BF16)
        dab = torch.empty((Bn, n), device=do.device, dtype=F32) if ctx.has_bias else None
        da = ops.gate_bwd(do, aux, gate, dgate, dab, rpb)
        dx = _w3_dgrad(ctx.link, da, wb)
        dw = _wgrad_gemm(da, a, [ctx.wparam]).view(ctx.wparam.shape)
        db = dab.sum(0) if ctx.has_bias else None
        return (dx, None, None, dgate, do, None if dmod is None else dmod[0],
                None if dmod is None else dmod[1], None, dw, db)


class LNModulateFn(Function):
    """adaLN: LN(x) * (1 + scale) + shift (Norm.py:16-23). x [B,T,d] bf16; shift/scale [B,d] bf16."""

    @staticmethod
    def forward(ctx, x, shift, scale):
        Bn, T, d = x.shape
        x2 = x.reshape(Bn * T, d)
        y, mean, rstd = ops.ln_modulate_fwd(x2, shift, scale, T)
        ctx.save_for_backward(x2, mean, rstd, scale)
        ctx.shape = (Bn, T, d)
        return y.view(Bn, T, d)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, scale = ctx.saved_tensors
        Bn, T, d = ctx.shape
        dmod = torch.empty((2, Bn, d), device=dy.device, dtype=BF16)     # written by the kernel
        dx = ops.ln_modulate_bwd(dy.reshape(Bn * T, d).contiguous(), x2, mean, rstd, scale, None,
                                 dmod[0], dmod[1], T)
        return dx.view(Bn, T, d), dmod[0], dmod[1]


class LNModulateResFn(Function):
    """adaLN modulate that also hands the residual stream through: returns (LN-mod(x), x).
    Using the second output as the residual makes x a single-use tensor for autograd, so the sum
    `d(LN path) + d(residual path)` happens inside the LN backward kernel (its `dres` operand)
    instead of a separate elementwise add."""

    @staticmethod
    def forward(ctx, x, shift, scale):
        Bn, T, d = x.shape
        x2 = x.reshape(Bn * T, d)
        y, mean, rstd = ops.ln_modulate_fwd(x2, shift, scale, T)
        ctx.save_for_backward(x2, mean, rstd, scale)
        ctx.shape = (Bn, T, d)
        ctx.set_materialize_grads(False)
        return y.view(Bn, T, d), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x2, mean, rstd, scale = ctx.saved_tensors
        Bn, T, d = ctx.shape
        if dy is None:      # only the residual path was used
            return dres, None, None
        dmod = torch.empty((2, Bn, d), device=dy.device, dtype=BF16)
        dr = None if dres is None else dres.reshape(Bn * T, d).contiguous()
        dx = ops.ln_modulate_bwd(dy.reshape(Bn * T, d).contiguous(), x2, mean, rstd, scale, dr,
                                 dmod[0], dmod[1], T)
        return dx.view(Bn, T, d), dmod[0], dmod[1]


class JointAttentionFn(Function):
    """QK-RMSNorm + 2-D RoPE (image tokens) + joint softmax attention over
    [image; text] (Attention.py:130-135,174-194,259-263,293,411-417).
    qkv_x [B*N,3d], qkv_c [B*M,3d]: packed raw projections (q | k | v)."""

    @staticmethod
    def forward(ctx, qkv_x, qkv_c, wq_x, wk_x, wq_c, wk_c, rope_cos, rope_sin, Bn, H, N, M):
        d = H * 64
        rope = (rope_cos, rope_sin) if rope_cos is not None else None
        qk_x = ops.qknorm_rope_fwd(qkv_x, wq_x, wk_x, rope, d, N)
        qk_c = ops.qknorm_rope_fwd(qkv_c, wq_c, wk_c, None, d, M)
        q = (qk_x[:, :d], qk_c[:, :d])
        k = (qk_x[:, d:], qk_c[:, d:])
        v = (qkv_x[:, 2 * d:], qkv_c[:, 2 * d:])
        # QK-RMSNorm bounds the logits: lets the kernel run a single-pass softmax
        bound = ops.qk_logit_bound(wq_x, wk_x, wq_c, wk_c, 0.125)
        o_x, o_c, lse = ops.attn_fwd(q, k, v, Bn, H, N, M, 0.125, logit_bound=bound)
        ctx.save_for_backward(qkv_x, qkv_c, qk_x, qk_c, o_x, o_c, lse, wq_x, wk_x, wq_c, wk_c,
                              rope_cos, rope_sin)
        ctx.dims = (Bn, H, N, M)
        ctx.set_materialize_grads(False)
        return o_x, o_c

    @staticmethod
    def backward(ctx, do_x, do_c):
        (qkv_x, qkv_c, qk_x, qk_c, o_x, o_c, lse, wq_x, wk_x, wq_c, wk_c, rope_cos,
         rope_sin) = ctx.saved_tensors
        Bn, H, N, M = ctx.dims
        d = H * 64
        do_x = torch.zeros_like(o_x) if do_x is None else do_x.contiguous()
        do_c = torch.zeros_like(o_c) if do_c is None else do_c.contiguous()
        dqkv_x, dqkv_c = torch.empty_like(qkv_x), torch.empty_like(qkv_c)
        dqk_x, dqk_c = torch.empty_like(qk_x), torch.empty_like(qk_c)
        # dq stays in the attention backward's fp32 accumulator: the QK-norm backward reads it from there
        # (no bf16 convert pass); only the k halves of dqk_x / dqk_c are ever written
        dq_acc = ops.attn_bwd((qk_x[:, :d], qk_c[:, :d]), (qk_x[:, d:], qk_c[:, d:]),
                              (qkv_x[:, 2 * d:], qkv_c[:, 2 * d:]), (o_x, o_c), lse, (do_x, do_c),
                              None, (dqk_x[:, d:], dqk_c[:, d:]),
                              (dqkv_x[:, 2 * d:], dqkv_c[:, 2 * d:]), Bn, H, N, M, 0.125)
        dw = zeropool.zeros((4, 64), qkv_x.device)
        rope = (rope_cos, rope_sin) if rope_cos is not None else None
        ops.qknorm_rope_bwd(dqk_x, qkv_x, wq_x, wk_x, rope, dqkv_x, dw[0], dw[1], d, N, dq_acc=dq_acc, acc_off=0)
        ops.qknorm_rope_bwd(dqk_c, qkv_c, wq_c, wk_c, None, dqkv_c, dw[2], dw[3], d, M, dq_acc=dq_acc, acc_off=N)
        return dqkv_x, dqkv_c, dw[0], dw[1], dw[2], dw[3], None, None, None, None, None, None


class QKVProjFn(Function):
    """Packed q|k|v projection (Attention.py:130-135) whose GEMM epilogue also emits the per-head
    RMSNorm * weight (+ 2-D RoPE on image tokens) of q and k (Attention.py:61-64,174-194).
    Returns (qkv raw [R,3d], qk normalised [R,2d]); qk is a non-differentiable side output --
    JointAttentionPreNormFn owns the backward of norm + RoPE and returns d(raw qkv)."""

    @staticmethod
    def forward(ctx, x, wb, wq, wk, cos, sin, tokens, *params):
        x2 = x.reshape(-1, x.shape[-1])
        d = wb.shape[0] // 3
        qk = torch.empty((x2.shape[0], 2 * d), device=x.device, dtype=BF16)
        qkv = ops.gemm(x2, wb, epilogue=ops.EPI_QKNORM, aux=qk,
                       qknorm=(wq.detach(), wk.detach(), cos, sin, tokens))
        ctx.save_for_backward(x2, wb)
        ctx.xshape = x.shape
        ctx.wparams = params
        ctx.mark_non_differentiable(qk)
        return qkv, qk

    @staticmethod
    def backward(ctx, dqkv, _dqk):
        x2, wb = ctx.saved_tensors
        dy2 = dqkv.reshape(-1, dqkv.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = ops.gemm(dy2, wb, b_major=1).reshape(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = _wgrad_gemm(dy2, x2, ctx.wparams)
        sizes = [p.shape[0] for p in ctx.wparams]
        grads = [g.view(p.shape) for g, p in zip(_split_rows(dw, sizes), ctx.wparams)]
        return (dx, None, None, None, None, None, None, *grads)


class JointAttentionPreNormFn(Function):
    """JointAttentionFn for projections that already carry their normalised / rotated q, k
    (QKVProjFn): same attention, same backward (which recomputes the norm from the raw qkv)."""

    @staticmethod
    def forward(ctx, qkv_x, qkv_c, qk_x, qk_c, wq_x, wk_x, wq_c, wk_c, rope_cos, rope_sin, Bn, H, N, M):
        d = H * 64
        q = (qk_x[:, :d], qk_c[:, :d])
        k = (qk_x[:, d:], qk_c[:, d:])
        v = (qkv_x[:, 2 * d:], qkv_c[:, 2 * d:])
        bound = ops.qk_logit_bound(wq_x, wk_x, wq_c, wk_c, 0.125)
        o_x, o_c, lse = ops.attn_fwd(q, k, v, Bn, H, N, M, 0.125, logit_bound=bound)
        ctx.save_for_backward(qkv_x, qkv_c, qk_x, qk_c, o_x, o_c, lse, wq_x, wk_x, wq_c, wk_c,
                              rope_cos, rope_sin)
        ctx.dims = (Bn, H, N, M)
        ctx.set_materialize_grads(False)
        return o_x, o_c

    @staticmethod
    def backward(ctx, do_x, do_c):
        grads = JointAttentionFn.backward(ctx, do_x, do_c)
        # JointAttentionFn: (dqkv_x, dqkv_c, dwq_x, dwk_x, dwq_c, dwk_c, 6 x None)
        return (grads[0], grads[1], None, None, grads[2], grads[3], grads[4], grads[5],
                None, None, None, None, None, None)


class TimestepEmbedFn(Function):
    """PositionalEncoding(t * time_scale) (PositionalEncoding.py:23-30, diff_model.py:306) -> bf16."""

    @staticmethod
    def forward(ctx, t, time_scale, denom):
        t = t.to(F32).contiguous()
        ctx.save_for_backward(t, time_scale, denom)
        return ops.timestep_embed_fwd(t, time_scale, denom)

    @staticmethod
    def backward(ctx, de):
        t, time_scale, denom = ctx.saved_tensors
        ds = zeropool.zeros(1, de.device)
        ops.timestep_embed_bwd(de.contiguous(), t, time_scale, denom, ds)
        return None, ds, None


class PatchifyFn(Function):
    """[B,C,H,W] -> bf16 tokens [B*N, C*p*p] (the im2col of the k=s=p conv,
    ImagePositionalEncoding.py:114-116,181-183)."""

    @staticmethod
    def forward(ctx, img, p):
        ctx.meta = (img.shape, img.dtype, p)
        return ops.patchify(img.contiguous(), p)

    @staticmethod
    def backward(ctx, g):
        (Bn, Cc, H, W), dtype, p = ctx.meta
        return ops.unpatchify(g.contiguous(), Bn, Cc, H, W, p, dtype), None


class UnpatchifyFn(Function):
    """tokens [B*N, C*p*p] bf16 -> [B,C,H,W] bf16 (patchify.py:41-72)."""

    @staticmethod
    def forward(ctx, tok, Bn, Cc, H, W, p):
        ctx.p = p
        return ops.unpatchify(tok.contiguous(), Bn, Cc, H, W, p, BF16)

    @staticmethod
    def backward(ctx, g):
        return ops.patchify(g.contiguous(), ctx.p), None, None, None, None, None


class RFLossFn(Function):
    """mean((v - (eps - x0))^2) in fp32 (model_trainer.py:429-446)."""

    @staticmethod
    def forward(ctx, v, eps, x0):
        loss, diff = ops.rf_loss_fwd(v.contiguous(), eps.contiguous(), x0.contiguous())
        ctx.save_for_backward(diff)
        ctx.dtype = v.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (diff,) = ctx.saved_tensors
        return ops.rf_loss_bwd(diff, g, ctx.dtype), None, None


def rf_loss(v, eps, x0):
    return RFLossFn.apply(v, eps, x0)


class SwiGLUHiddenFn(Function):
    """a = silu(x1) * x2 with [x1 | x2] = x @ W12^T + b12 (xformers SwiGLU up to w3, MLP.py:19,32).
    One Function so that the bias gradient of w12 falls out of the activation-backward kernel
    instead of a second pass over dh12."""

    @staticmethod
    def forward(ctx, x, wb, w12, b12):
        x2 = x.reshape(-1, x.shape[-1])
        bias = None if b12 is None else b12.detach()
        if FUSED_SWIGLU and x2.shape[0] > 128 and wb.shape[0] % 256 == 0 and (bias is None or bias.dtype == F32):
            # activation in the GEMM epilogue: the [R, 8d] pre-activation is written once (for the
            # backward) and never re-read by a separate activation kernel
            # (inference: nobody reads the pre-activation -- the epilogue skips its two stores per chunk,
            #  2/3 of this GEMM's output bytes)
            h12 = (torch.empty((x2.shape[0], wb.shape[0]), device=x.device, dtype=BF16)
                   if any(ctx.needs_input_grad) else None)
            a = ops.gemm(x2, wb, bias=bias, epilogue=ops.EPI_SWIGLU, aux=h12)
        else:
            h12 = ops.gemm(x2, wb, bias=bias)
            a = ops.swiglu_fwd(h12)
        ctx.save_for_backward(x2, wb, h12)
        ctx.wparam = w12
        ctx.meta = (x.shape, b12 is not None)
        ctx.link = None
        if (FUSED_SWIGLU_BWD and any(ctx.needs_input_grad) and a.is_contiguous()
                and ops.swiglu_bwd_fusable(a.shape[0], a.shape[1])):
            if len(_SWIGLU_LINKS) > 8:
                _SWIGLU_LINKS.clear()     # activations nobody claimed (hidden() used on its own)
            ctx.link = _SWIGLU_LINKS[a.data_ptr()] = _SwigluLink(h12, b12 is not None)
        return a.reshape(*x.shape[:-1], a.shape[-1])

    @staticmethod
    def backward(ctx, da):
        x2, wb, h12 = ctx.saved_tensors
        xshape, has_bias = ctx.meta
        link = ctx.link
        if link is not None and link.paired:
            # the w3 node already applied the SwiGLU backward in its GEMM epilogue; `da` is its dummy view
            if link.dh12 is None or da.data_ptr() != link.dummy_ptr:
                raise RuntimeError("fused SwiGLU backward: the activation must feed exactly one w3 GEMM")
            dh, db = link.dh12, link.db
            link.dh12 = link.db = link.h12 = None
        else:
            db = zeropool.zeros(h12.shape[1], da.device) if has_bias else None
            dh = ops.swiglu_bwd(da.reshape(-1, da.shape[-1]).contiguous(), h12, db)
        dx = ops.gemm(dh, wb, b_major=1).reshape(xshape)
        dw = _wgrad_gemm(dh, x2, [ctx.wparam])
        return dx, None, dw, db


class TextFrontFn(Function):
    """The whole text front-end (diff_model.py:323-326) as one node:
    c' = cat[c_proj(s1 * RMSNorm(c[:, :77])), c_proj2(s2 * RMSNorm(c[:, 77:]))].
    Keeping it one Function lets the dgrad that feeds the two scalar / two norm-weight gradients
    stay in fp32 (those are sums over B*77*2304 terms with heavy cancellation)."""

    @staticmethod
    def forward(ctx, c, w1, w2, s1, s2, split, wb1, wb2, pw1, pw2):
        c = c.contiguous()
        Bn, M, _ = c.shape
        n1, n2, rstd = ops.text_norm_fwd(c, w1, w2, s1, s2, split)
        t1, t2, d = split, M - split, wb1.shape[0]
        out = torch.empty((Bn * M, d), device=c.device, dtype=BF16)
        ops.gemm(n1, wb1, out=out, remap=(t1, M, 0))
        if t2:
            ops.gemm(n2, wb2, out=out, remap=(t2, M, t1))
        ctx.save_for_backward(c, rstd, w1, w2, s1, s2, n1, n2, wb1, wb2)
        ctx.dims = (Bn, t1, t2, d)
        return out.view(Bn, M, d)

    @staticmethod
    def backward(ctx, g):
        c, rstd, w1, w2, s1, s2, n1, n2, wb1, wb2 = ctx.saved_tensors
        Bn, t1, t2, d = ctx.dims
        dw1, dw2 = zeropool.zeros(w1.shape, w1.device), zeropool.zeros(w2.shape, w2.device)
        ds1, ds2 = zeropool.zeros(s1.shape, s1.device), zeropool.zeros(s2.shape, s2.device)
        g1 = g[:, :t1].reshape(Bn * t1, d)
        dn1 = ops.gemm(g1, wb1, b_major=1, out_dtype=F32)
        dpw1 = ops.gemm(g1, n1, a_major=1, b_major=1, out_dtype=F32)
        ops.text_norm_bwd(dn1, c, rstd, w1, s1, dw1, ds1, 0, t1)
        dpw2 = None
        if t2:
            g2 = g[:, t1:].reshape(Bn * t2, d)
            dn2 = ops.gemm(g2, wb2, b_major=1, out_dtype=F32)
            dpw2 = ops.gemm(g2, n2, a_major=1, b_major=1, out_dtype=F32)
            ops.text_norm_bwd(dn2, c, rstd, w2, s2, dw2, ds2, t1, t2)
        return None, dw1, dw2, ds1, ds2, None, None, None, dpw1, dpw2

"""Raw (non-autograd) Python wrappers over the C-ABI kernels.

Every function validates shapes/dtypes in Python (the reference's error
convention is Python exceptions), launches on the current torch stream and
returns torch tensors allocated by PyTorch's caching allocator.  No function
here has a CPU or eager fallback.
"""
import ctypes as C
import os

import torch

from . import _lib, zeropool
from ._lib import (EPI_GATE_RESID, EPI_NONE, EPI_QKNORM, EPI_RESID, EPI_SILU, EPI_SWIGLU,  # noqa: F401
                   EPI_SWIGLU_BWD)

BF16 = torch.bfloat16
F32 = torch.float32
LN_EPS = 1e-5                      # nn.LayerNorm default (Norm.py:10)
RMS_EPS = 1.1920928955078125e-07   # nn.RMSNorm(eps=None) -> finfo(float32).eps (torch 2.11)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmdit ops need CUDA tensors on a B200; there is no CPU fallback")


def _p(t):
    return 0 if t is None else t.data_ptr()


def _s():
    if not torch.cuda.is_available():
        raise RuntimeError("mmdit ops need CUDA tensors on a B200; there is no CPU fallback")
    return torch.cuda.current_stream().cuda_stream


def _workspace(nfloats, device):
    """Scratch fp32 buffer from PyTorch's caching allocator (the library never allocates)."""
    return torch.empty(int(nfloats), device=device, dtype=F32)


def _plan_split(M, N, K, sms=148):
    """Split-K factor for an fp32-output GEMM: minimise waves * (k-blocks per unit + epilogue)."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    kb = (K + 63) // 64
    if tiles >= sms or kb < 16:
        return 1
    best, best_s = None, 1
    for s in range(1, min(kb // 4, 32) + 1):
        per = -(-kb // s)
        units = tiles * (-(-kb // per))
        t = -(-units // sms) * (per + 6.0) + (0.0 if s == 1 else 2.0 + 0.5 * s)
        if best is None or t < best - 1e-9:
            best, best_s = t, s
    return best_s


def _rowmajor2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} {t.stride()}")


# --------------------------------------------------------------------- GEMM
def gemm(A, B, **kw):
    """D[M,N] = epilogue(A @ B^T).  A is stored [M,K] (a_major=0) or [K,M] (a_major=1);
    B is stored [N,K] (b_major=0) or [K,N] (b_major=1).  Keyword arguments: see _gemm."""
    return _gemm(A, B, **kw)


def _gemm(A, B, *, a_major=0, b_major=0, out=None, out_dtype=BF16, accumulate=False, split_k=0,
          epilogue=EPI_NONE, bias=None, gate=None, rows_per_gate=0, resid=None, aux=None,
          remap=None, out_rows=None, force_block_n=0, simt=False, _logical_m=None, qknorm=None):
    """Body of gemm(); the split-K plans below recurse into _gemm so that a wrapper installed on
    ops.gemm (bench.py's per-launch timers) sees every logical GEMM exactly once."""
    _need_cuda(A, B)
    _rowmajor2d(A, "A")
    _rowmajor2d(B, "B")
    if A.dtype != BF16 or B.dtype != BF16:
        raise TypeError("gemm operands must be bfloat16")
    M, K = (A.shape[1], A.shape[0]) if a_major else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_major else (B.shape[0], B.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: reduction dims differ ({K} vs {Kb})")
    if (out is None and out_dtype == BF16 and M <= 128 and K >= 2048 and epilogue == EPI_NONE
            and bias is None and aux is None and remap is None and not simt):
        # few output tiles, long reduction (adaLN dgrads with M = batch): split-K over all SMs
        # into an fp32 buffer, then one cast
        acc = _gemm(A, B, a_major=a_major, b_major=b_major, out_dtype=F32, accumulate=True)
        return acc.to(BF16)
    dst = None
    if (out is not None and out.dtype == F32 and not accumulate and split_k == 0 and out.is_contiguous()
            and tuple(out.shape) == (M, N)):
        dst, out = out, None     # caller-provided destination (e.g. a gradient-bucket slot): still free to split
    if (out is None and (out_dtype == F32 or dst is not None) and not accumulate and remap is None
            and epilogue == EPI_NONE
            and bias is None and aux is None and not simt and N % 8 == 0 and split_k == 0):
        # wgrad-style outputs (few tiles, long reduction): split-K so that every SM has work; each
        # split writes its own fp32 slice (plain coalesced stores) and one small kernel folds them
        s = _plan_split(M, N, K)
        if s > 1:
            ws = torch.empty((s, M, N), device=A.device, dtype=F32)
            _gemm(A, B, a_major=a_major, b_major=b_major, out=ws.view(s * M, N), split_k=s,
                  force_block_n=force_block_n, _logical_m=M)
            out = dst if dst is not None else torch.empty((M, N), device=A.device, dtype=F32)
            _lib.check(_lib.lib().mmdit_fold_slices_f32(_p(ws), _p(out), M * N, s, M * N, 0, _s()),
                       "mmdit_fold_slices_f32")
            return out
    if dst is not None:
        out = dst
    if out is None:
        rows = out_rows if out_rows is not None else M
        if remap is not None and out_rows is None:
            raise ValueError("gemm: remap needs out_rows or out")
        out = (torch.zeros if (accumulate or remap is not None) else torch.empty)(
            (rows, N // 2 if epilogue == EPI_SWIGLU else N), device=A.device, dtype=out_dtype)
    _rowmajor2d(out, "out")
    a = _lib.GemmArgs()
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), out.data_ptr()
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldb, a.ldd = A.stride(0), B.stride(0), out.stride(0)
    a.a_major, a.b_major = int(a_major), int(b_major)
    a.d_fp32 = int(out.dtype == F32)
    a.accumulate, a.split_k, a.epilogue = int(accumulate), int(split_k), int(epilogue)
    if bias is not None:
        a.bias, a.bias_fp32 = bias.data_ptr(), int(bias.dtype == F32)
    if gate is not None:
        _rowmajor2d(gate, "gate")
        a.gate, a.rows_per_gate, a.ld_gate = gate.data_ptr(), int(rows_per_gate), gate.stride(0)
    if resid is not None:
        _rowmajor2d(resid, "resid")
        a.resid, a.ldr = resid.data_ptr(), resid.stride(0)
    if aux is not None:
        _rowmajor2d(aux, "aux")
        a.aux, a.ld_aux = aux.data_ptr(), aux.stride(0)
    if remap is not None:
        a.remap_rows, a.remap_batch_rows, a.remap_offset = (int(x) for x in remap)
    a.force_block_n = int(force_block_n)
    if epilogue == EPI_QKNORM:
        # qknorm = (wq fp32[64], wk fp32[64], rope_cos | None, rope_sin | None, tokens per sample);
        # aux receives the normalised / rotated q | k ([M, 2N/3])
        if qknorm is None or aux is None:
            raise ValueError("gemm: the QK-norm epilogue needs qknorm=(wq, wk, cos, sin, tokens) and aux")
        wq_, wk_, cos_, sin_, tokens_ = qknorm
        a.qk_wq, a.qk_wk = wq_.data_ptr(), wk_.data_ptr()
        a.rope_cos = cos_.data_ptr() if cos_ is not None else None
        a.rope_sin = sin_.data_ptr() if sin_ is not None else None
        a.qk_tokens, a.qk_eps = int(tokens_), RMS_EPS
    L = _lib.lib()
    fn = L.mmdit_gemm_bf16_simt if simt else L.mmdit_gemm_bf16
    _lib.check(fn(C.byref(a), _s()), "mmdit_gemm_bf16")
    return out


def swiglu_bwd_fusable(rows, hidden):
    """Shapes the fused w3-dgrad + SwiGLU-backward GEMM accepts (whole 128-row / 256-column tiles)."""
    return rows > 128 and rows % 128 == 0 and hidden % 256 == 0


def gemm_swiglu_bwd(dy, w3b, h12, db12, debug=0):
    """dh12 [R, 2*hidden] = SwiGLU backward of (dy @ W3) against the saved pre-activations h12 = [x1 | x2],
    in the epilogue of the data-gradient GEMM of w3 (w3b: bf16 [d, hidden]); the activation gradient is
    never materialised.  db12 (fp32 [2*hidden], zero-initialised by the caller, or None) receives the
    column sums of dh12 (the w12 bias gradient).  Bit-identical to gemm(b_major=1) + swiglu_bwd."""
    _need_cuda(dy, w3b, h12)
    _rowmajor2d(dy, "dy")
    R, d = dy.shape
    hid = w3b.shape[1]
    assert w3b.shape[0] == d and h12.shape == (R, 2 * hid) and h12.is_contiguous() and swiglu_bwd_fusable(R, hid)
    dh12 = torch.empty_like(h12)
    a = _lib.GemmArgs()
    a.A, a.B, a.D = dy.data_ptr(), w3b.data_ptr(), dh12.data_ptr()
    a.M, a.N, a.K = R, hid, d
    a.lda, a.ldb, a.ldd = dy.stride(0), w3b.stride(0), dh12.stride(0)
    a.a_major, a.b_major = 0, 1
    a.epilogue = EPI_SWIGLU_BWD
    a.reserved = int(debug)
    a.aux, a.ld_aux = h12.data_ptr(), h12.stride(0)
    part = None
    if db12 is not None:
        part = torch.empty((R // 32, 2 * hid), device=dy.device, dtype=F32)
        a.colsum_partial = part.data_ptr()
    _lib.check(_lib.lib().mmdit_gemm_bf16(C.byref(a), _s()), "mmdit_gemm_bf16 (SwiGLU backward)")
    if part is not None:
        _lib.check(_lib.lib().mmdit_fold_rows_f32(_p(part), _p(db12), R // 32, 2 * hid, 2 * hid, _s()),
                   "mmdit_fold_rows_f32")
    return dh12


# ---------------------------------------------------------------- attention
def _fill_streams(args, name_ptr, name_ld, tensors):
    for s, t in enumerate(tensors):
        if t is None:
            continue
        _rowmajor2d(t, name_ptr)
        getattr(args, name_ptr)[s] = t.data_ptr()
        getattr(args, name_ld)[s] = t.stride(0)


def qk_logit_bound(wq_x, wk_x, wq_c, wk_c, scale):
    """Device scalar bounding |scale * q.k| after per-head RMSNorm with these norm weights."""
    out = torch.empty(1, device=wq_x.device, dtype=F32)
    _lib.check(_lib.lib().mmdit_qk_logit_bound(_p(wq_x), _p(wk_x), _p(wq_c), _p(wk_c), float(scale), _p(out),
                                               _s()), "mmdit_qk_logit_bound")
    return out


def attn_fwd(q, k, v, B, H, N, M, scale, logit_bound=None):
    """q/k/v: pairs (image, text) of [B*rows, H*64] bf16 views (unit inner stride).
    Returns (o_x [B*N,H*64], o_c [B*M,H*64] or None, lse [B,H,N+M])."""
    _need_cuda(q[0])
    dev, d = q[0].device, H * 64
    o = [torch.empty((B * N, d), device=dev, dtype=BF16),
         torch.empty((B * M, d), device=dev, dtype=BF16) if M > 0 else None]
    lse = torch.empty((B, H, N + M), device=dev, dtype=F32)
    a = _lib.AttnArgs()
    _fill_streams(a, "q", "ld_q", q)
    _fill_streams(a, "k", "ld_k", k)
    _fill_streams(a, "v", "ld_v", v)
    _fill_streams(a, "o", "ld_o", o)
    a.lse = lse.data_ptr()
    a.logit_bound = _p(logit_bound) or None
    a.B, a.H, a.N, a.M, a.head_dim, a.scale = B, H, N, M, 64, float(scale)
    _lib.check(_lib.lib().mmdit_attn_fwd(C.byref(a), _s()), "mmdit_attn_fwd")
    return o[0], o[1], lse


def attn_bwd(q, k, v, o, lse, d_o, dq, dk, dv, B, H, N, M, scale):
    """Writes dq/dk/dv (pairs of [B*rows, H*64] bf16 views) for the joint attention.
    dq=None: no bf16 copy of dq is written; the fp32 accumulator [B, N+M, H*64] (image rows first in
    every sample) is returned for qknorm_rope_bwd(..., dq_acc=...) to consume."""
    _need_cuda(q[0])
    dev = q[0].device
    T = N + M
    delta = torch.empty((B, H, T), device=dev, dtype=F32)
    dq_acc = torch.empty((B, T, H * 64), device=dev, dtype=F32)
    a = _lib.AttnArgs()
    _fill_streams(a, "q", "ld_q", q)
    _fill_streams(a, "k", "ld_k", k)
    _fill_streams(a, "v", "ld_v", v)
    _fill_streams(a, "o", "ld_o", o)
    _fill_streams(a, "d_o", "ld_do", d_o)
    if dq is not None:
        _fill_streams(a, "dq", "ld_dq", dq)
    _fill_streams(a, "dk", "ld_dk", dk)
    _fill_streams(a, "dv", "ld_dv", dv)
    a.lse, a.delta, a.dq_acc = lse.data_ptr(), delta.data_ptr(), dq_acc.data_ptr()
    a.B, a.H, a.N, a.M, a.head_dim, a.scale = B, H, N, M, 64, float(scale)
    _lib.check(_lib.lib().mmdit_attn_bwd(C.byref(a), _s()), "mmdit_attn_bwd")
    return dq_acc if dq is None else None


# ------------------------------------------------------------- row kernels
_row_generation = 1 if os.environ.get("MMDIT_ROW_KERNELS") == "1" else 2   # mirrors the library's initial value


def set_row_kernel_generation(generation):
    """2 (default): csrc/rowwise2.cu + csrc/qknorm2.cu; 1: the first-generation kernels (A/B and regression knob)."""
    global _row_generation
    _lib.check(_lib.lib().mmdit_set_row_kernel_generation(int(generation)), "mmdit_set_row_kernel_generation")
    _row_generation = int(generation)


def fused_gate_ln_max_columns():
    """Widest row the one-pass gated residual + LayerNorm-modulate kernel is worth using for: the first
    generation falls to one block per SM above 1024 columns (106 us vs 30 + 31 us at d = 1536); the second
    runs 38 us there against 30 + 21 us for the two kernels."""
    return 1536 if _row_generation >= 2 else 1024


def ln_modulate_fwd(x, shift, scale, rows_per_batch, save_stats=True):
    """x [R,d] bf16; shift/scale [B,d] bf16 views (same row stride)."""
    _need_cuda(x)
    R, d = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(R, device=x.device, dtype=F32) if save_stats else None
    rstd = torch.empty(R, device=x.device, dtype=F32) if save_stats else None
    assert shift.stride(0) == scale.stride(0) and x.is_contiguous()
    _lib.check(_lib.lib().mmdit_ln_modulate_fwd(
        _p(x), _p(shift), _p(scale), _p(y), _p(mean), _p(rstd), R, d, rows_per_batch,
        shift.stride(0), LN_EPS, _s()), "mmdit_ln_modulate_fwd")
    return y, mean, rstd


def ln_modulate_bwd(dy, x, mean, rstd, scale, dres, dshift, dscale, rows_per_batch):
    """dshift/dscale: fp32 [B,d] views that are accumulated into. Returns dx bf16."""
    R, d = x.shape
    dx = torch.empty_like(x)
    assert dy.is_contiguous() and x.is_contiguous() and (dres is None or dres.is_contiguous())
    assert dshift.stride(0) == dscale.stride(0)
    ws = _workspace(_lib.lib().mmdit_rowreduce_workspace_floats(R, d, rows_per_batch), x.device)
    _lib.check(_lib.lib().mmdit_ln_modulate_bwd(
        _p(dy), _p(x), _p(mean), _p(rstd), _p(scale), _p(dres), _p(dx), _p(dshift), _p(dscale),
        int(dshift.dtype == BF16), _p(ws), R, d, rows_per_batch, scale.stride(0), dshift.stride(0), _s()), "mmdit_ln_modulate_bwd")
    return dx


def gate_bwd(dout, a, gate, dgate, dab, rows_per_batch):
    R, d = dout.shape
    da = torch.empty_like(dout)
    assert dout.is_contiguous() and a.is_contiguous()
    ws = _workspace(_lib.lib().mmdit_rowreduce_workspace_floats(R, d, rows_per_batch), dout.device)
    _lib.check(_lib.lib().mmdit_gate_bwd(
        _p(dout), _p(a), _p(gate), _p(da), _p(dgate), int(dgate.dtype == BF16), _p(dab), _p(ws), R, d,
        rows_per_batch,
        gate.stride(0), dgate.stride(0), 0 if dab is None else dab.stride(0), _s()),
        "mmdit_gate_bwd")
    return da


def text_norm_fwd(c, w1, w2, s1, s2, split):
    """c [B,M,dt] bf16 -> (out1 [B*split,dt], out2 [B*(M-split),dt], rstd [B*M])."""
    _need_cuda(c)
    Bn, M, dt = c.shape
    assert c.is_contiguous() and c.dtype == BF16
    out1 = torch.empty((Bn * split, dt), device=c.device, dtype=BF16)
    out2 = torch.empty((Bn * (M - split), dt), device=c.device, dtype=BF16) if M > split else None
    rstd = torch.empty(Bn * M, device=c.device, dtype=F32)
    _lib.check(_lib.lib().mmdit_text_norm_fwd(
        _p(c), _p(w1), _p(w2), _p(s1), _p(s2), _p(out1), _p(out2), _p(rstd), Bn, M, split, dt,
        RMS_EPS, _s()), "mmdit_text_norm_fwd")
    return out1, out2, rstd


def text_norm_bwd(dn, c, rstd, w, sigma, dw, dsigma, tok0, ntok):
    Bn, M, dt = c.shape
    assert dn.is_contiguous()
    _lib.check(_lib.lib().mmdit_text_norm_bwd(
        _p(dn), int(dn.dtype == F32), _p(c), _p(rstd), _p(w), _p(sigma), _p(dw), _p(dsigma), Bn, M, tok0,
        ntok, dt,
        _s()), "mmdit_text_norm_bwd")


def qknorm_rope_fwd(qkv, wq, wk, rope, d, tokens_per_sample):
    """qkv [R, >=2d] raw projections (q at col 0, k at col d). Returns [R, 2d] bf16."""
    _need_cuda(qkv)
    R = qkv.shape[0]
    out = torch.empty((R, 2 * d), device=qkv.device, dtype=BF16)
    cos, sin = rope if rope is not None else (None, None)
    _lib.check(_lib.lib().mmdit_qknorm_rope_fwd(
        _p(qkv), _p(wq), _p(wk), _p(cos), _p(sin), _p(out), R, d, qkv.stride(0), out.stride(0),
        tokens_per_sample, RMS_EPS, _s()), "mmdit_qknorm_rope_fwd")
    return out


def qknorm_rope_bwd(dqk, qkv, wq, wk, rope, dqkv, dwq, dwk, d, tokens_per_sample, dq_acc=None, acc_off=0):
    """dq_acc (fp32 [B, T, d] from attn_bwd(dq=None)): the q half of the gradient is read from it (this
    stream's rows start at acc_off in every sample); only the k half of dqk is read then."""
    R = qkv.shape[0]
    cos, sin = rope if rope is not None else (None, None)
    if dq_acc is not None:
        _lib.check(_lib.lib().mmdit_qknorm_rope_bwd_acc(
            _p(dq_acc), dq_acc.shape[1], acc_off, _p(dqk), _p(qkv), _p(wq), _p(wk), _p(cos), _p(sin), _p(dqkv),
            _p(dwq), _p(dwk), R, d, dqk.stride(0), qkv.stride(0), dqkv.stride(0), tokens_per_sample, RMS_EPS,
            _s()), "mmdit_qknorm_rope_bwd_acc")
        return
    _lib.check(_lib.lib().mmdit_qknorm_rope_bwd(
        _p(dqk), _p(qkv), _p(wq), _p(wk), _p(cos), _p(sin), _p(dqkv), _p(dwq), _p(dwk), R, d,
        dqk.stride(0), qkv.stride(0), dqkv.stride(0), tokens_per_sample, RMS_EPS, _s()),
        "mmdit_qknorm_rope_bwd")


def swiglu_fwd(h12):
    R, two_h = h12.shape
    assert h12.is_contiguous()
    a = torch.empty((R, two_h // 2), device=h12.device, dtype=BF16)
    _lib.check(_lib.lib().mmdit_swiglu_fwd(_p(h12), _p(a), R, two_h // 2, _s()), "mmdit_swiglu_fwd")
    return a


def swiglu_bwd(da, h12, db12):
    R, two_h = h12.shape
    assert da.is_contiguous() and h12.is_contiguous()
    dh12 = torch.empty_like(h12)
    ws = None
    if db12 is not None:
        ws = _workspace(_lib.lib().mmdit_swiglu_bwd_workspace_floats(R, two_h // 2), h12.device)
    _lib.check(_lib.lib().mmdit_swiglu_bwd(_p(da), _p(h12), _p(dh12), _p(db12), _p(ws), R, two_h // 2, _s()),
               "mmdit_swiglu_bwd")
    return dh12


def timestep_embed_fwd(t, time_scale, denom):
    Bn, d = t.shape[0], denom.shape[0]
    out = torch.empty((Bn, d), device=t.device, dtype=BF16)
    _lib.check(_lib.lib().mmdit_timestep_embed_fwd(_p(t), _p(time_scale), _p(denom), _p(out), Bn, d,
                                                   _s()), "mmdit_timestep_embed_fwd")
    return out


def timestep_embed_bwd(de, t, time_scale, denom, dscale):
    Bn, d = de.shape
    assert de.is_contiguous()
    _lib.check(_lib.lib().mmdit_timestep_embed_bwd(_p(de), _p(t), _p(time_scale), _p(denom),
                                                   _p(dscale), Bn, d, _s()),
               "mmdit_timestep_embed_bwd")


def patchify(img, p):
    _need_cuda(img)
    Bn, Cc, H, W = img.shape
    assert img.is_contiguous() and img.dtype in (F32, BF16)
    tok = torch.empty((Bn * (H // p) * (W // p), Cc * p * p), device=img.device, dtype=BF16)
    _lib.check(_lib.lib().mmdit_patchify(_p(img), int(img.dtype == F32), _p(tok), Bn, Cc, H, W, p,
                                         _s()), "mmdit_patchify")
    return tok


def unpatchify(tok, Bn, Cc, H, W, p, dtype=BF16):
    assert tok.is_contiguous() and tok.dtype == BF16
    img = torch.empty((Bn, Cc, H, W), device=tok.device, dtype=dtype)
    _lib.check(_lib.lib().mmdit_unpatchify(_p(tok), _p(img), int(dtype == F32), Bn, Cc, H, W, p,
                                           _s()), "mmdit_unpatchify")
    return img


def rf_noise(x0, eps, t):
    _need_cuda(x0)
    assert x0.dtype == eps.dtype and x0.dtype in (F32, BF16) and x0.is_contiguous()
    xt = torch.empty(x0.shape, device=x0.device, dtype=F32)
    _lib.check(_lib.lib().mmdit_rf_noise(_p(x0), _p(eps), int(x0.dtype == F32), _p(t), _p(xt),
                                         x0.shape[0], x0[0].numel(), _s()), "mmdit_rf_noise")
    return xt


def rf_loss_fwd(v, eps, x0):
    assert v.is_contiguous() and eps.is_contiguous() and x0.is_contiguous()
    assert eps.dtype == x0.dtype
    diff = torch.empty(v.shape, device=v.device, dtype=F32)
    loss = torch.empty((), device=v.device, dtype=F32)
    _lib.check(_lib.lib().mmdit_rf_loss_fwd(_p(v), int(v.dtype == F32), _p(eps), _p(x0),
                                            int(x0.dtype == F32), _p(diff), _p(loss), v.numel(),
                                            _s()), "mmdit_rf_loss_fwd")
    return loss, diff


def rf_loss_bwd(diff, upstream, dtype):
    dv = torch.empty(diff.shape, device=diff.device, dtype=dtype)
    up = upstream.to(F32).reshape(1).contiguous()
    _lib.check(_lib.lib().mmdit_rf_loss_bwd(_p(diff), _p(up), _p(dv), int(dtype == F32),
                                            diff.numel(), _s()), "mmdit_rf_loss_bwd")
    return dv


def cfg_euler_step(x, v, cfg_scale, dt):
    """x fp32 [B,...] updated in place from v [2B,...]."""
    assert x.dtype == F32 and x.is_contiguous() and v.is_contiguous()
    _lib.check(_lib.lib().mmdit_cfg_euler_step(_p(x), _p(v), int(v.dtype == F32), x.numel(),
                                               float(cfg_scale), float(dt), _s()),
               "mmdit_cfg_euler_step")
    return x


def colsum(x, out=None):
    R, n = x.shape
    if out is None:
        out = zeropool.zeros(n, x.device)
    _lib.check(_lib.lib().mmdit_colsum_bf16(_p(x), _p(out), R, n, x.stride(0), _s()),
               "mmdit_colsum_bf16")
    return out


def fold_rows(x, out):
    R, n = x.shape
    _lib.check(_lib.lib().mmdit_fold_rows_f32(_p(x), _p(out), R, n, x.stride(0), _s()),
               "mmdit_fold_rows_f32")
    return out


def cast_bf16(x, out=None):
    assert x.dtype == F32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.check(_lib.lib().mmdit_cast_f32_bf16(_p(x), _p(out), x.numel(), _s()),
               "mmdit_cast_f32_bf16")
    return out


def gate_residual_fwd(a, gate, resid, rows_per_batch):
    """out = a * gate[row // rows_per_batch] + resid (all bf16, a/resid [R,d] contiguous)."""
    R, d = a.shape
    assert a.is_contiguous() and resid.is_contiguous()
    out = torch.empty_like(a)
    _lib.check(_lib.lib().mmdit_gate_residual_fwd(_p(a), _p(gate), _p(resid), _p(out), R, d,
                                                  rows_per_batch, gate.stride(0), _s()),
               "mmdit_gate_residual_fwd")
    return out


def gate_residual_ln_fwd(a, gate, resid, shift, scale, rows_per_batch):
    """x' = a * gate[b] + resid and y = LN(x') * (1 + scale[b]) + shift[b] in one pass.
    Returns (x' bf16 [R,d], y bf16 [R,d], mean fp32 [R], rstd fp32 [R])."""
    R, d = a.shape
    assert a.is_contiguous() and resid.is_contiguous() and shift.stride(0) == scale.stride(0)
    xo, y = torch.empty_like(a), torch.empty_like(a)
    mean = torch.empty(R, device=a.device, dtype=F32)
    rstd = torch.empty(R, device=a.device, dtype=F32)
    _lib.check(_lib.lib().mmdit_gate_residual_ln_fwd(
        _p(a), _p(gate), _p(resid), _p(shift), _p(scale), _p(xo), _p(y), _p(mean), _p(rstd), R, d,
        rows_per_batch, gate.stride(0), shift.stride(0), LN_EPS, _s()), "mmdit_gate_residual_ln_fwd")
    return xo, y, mean, rstd

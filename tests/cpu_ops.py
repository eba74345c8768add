"""TEST-ONLY stand-ins for mmdit.ops: the same call signatures and in/out conventions, computed with
plain torch on whatever device the tensors live on (fp32 math, bf16 where the kernels round).
Installed over `mmdit.ops` by tests/test_model_flow_cpu.py so that the product's HOST logic
(reference-shaped modules, autograd glue, gradient buckets) can be executed and compared with the
oracle without a GPU.  Nothing outside tests/ may import this file: the product has no CPU path.
The formulas are the torch restatements tools/kernel_probe.py checks the CUDA kernels against."""
import torch
import torch.nn.functional as F

BF16, F32 = torch.bfloat16, torch.float32
LN_EPS = 1e-5
RMS_EPS = 1.1920928955078125e-07
EPI_NONE, EPI_GATE_RESID, EPI_SILU, EPI_RESID, EPI_SWIGLU, EPI_QKNORM = 0, 1, 2, 3, 4, 5


def _bf(x):
    return x.to(BF16)


def _rows(t, rpb):
    return t.float().repeat_interleave(rpb, 0)


# ------------------------------------------------------------------ q/k norm + rope
def _qknorm(x, w, H, rope, tokens):
    """x [R, H*64] fp32 -> per-head RMSNorm * w (bf16-rounded), then interleaved-pair rotation."""
    R = x.shape[0]
    n = F.rms_norm(x.view(R, H, 64), (64,), w.float(), RMS_EPS)
    n = n + (n.to(BF16).float() - n).detach()
    if rope is None or rope[0] is None:
        return n.reshape(R, H * 64)
    cos, sin = rope
    tok = torch.arange(R, device=x.device) % tokens
    c = cos[tok].repeat_interleave(2, -1)[:, None, :]
    s = sin[tok].repeat_interleave(2, -1)[:, None, :]
    x1, x2 = n[..., 0::2], n[..., 1::2]
    rot = torch.stack((-x2, x1), -1).reshape(R, H, 64)
    return (n * c + rot * s).reshape(R, H * 64)


def qknorm_rope_fwd(qkv, wq, wk, rope, d, tokens_per_sample):
    H = d // 64
    q = _qknorm(qkv[:, :d].float(), wq, H, rope, tokens_per_sample)
    k = _qknorm(qkv[:, d:2 * d].float(), wk, H, rope, tokens_per_sample)
    return _bf(torch.cat([q, k], 1))


def qknorm_rope_bwd(dqk, qkv, wq, wk, rope, dqkv, dwq, dwk, d, tokens_per_sample, dq_acc=None, acc_off=0):
    H = d // 64
    if dq_acc is not None:      # q half from the fp32 accumulator [B, T, d], rounded to bf16 like the old copy
        Bn = qkv.shape[0] // tokens_per_sample
        gq_in = _bf(dq_acc[:, acc_off:acc_off + tokens_per_sample].reshape(Bn * tokens_per_sample, d))
        dqk = torch.cat([gq_in, dqk[:, d:]], 1)
    with torch.enable_grad():
        q = qkv[:, :d].float().detach().requires_grad_(True)
        k = qkv[:, d:2 * d].float().detach().requires_grad_(True)
        a, b = wq.detach().float().requires_grad_(True), wk.detach().float().requires_grad_(True)
        out = (_qknorm(q, a, H, rope, tokens_per_sample) * dqk[:, :d].float()).sum() + \
              (_qknorm(k, b, H, rope, tokens_per_sample) * dqk[:, d:].float()).sum()
        gq, gk, ga, gb = torch.autograd.grad(out, (q, k, a, b))
    dqkv[:, :d] = _bf(gq)
    dqkv[:, d:2 * d] = _bf(gk)
    dwq += ga
    dwk += gb


# ------------------------------------------------------------------------- GEMM
def gemm(A, B, *, a_major=0, b_major=0, out=None, out_dtype=BF16, accumulate=False, split_k=0,
         epilogue=EPI_NONE, bias=None, gate=None, rows_per_gate=0, resid=None, aux=None, remap=None,
         out_rows=None, force_block_n=0, simt=False, _logical_m=None, qknorm=None):
    assert A.dtype == BF16 and B.dtype == BF16
    A2 = A.float().t() if a_major else A.float()
    B2 = B.float() if b_major else B.float().t()
    acc = A2 @ B2
    if bias is not None:
        acc = acc + bias.float()
    M, N = acc.shape
    if epilogue in (EPI_SILU, EPI_GATE_RESID, EPI_SWIGLU) and aux is not None:
        aux.copy_(acc)
    if epilogue == EPI_SILU:
        res = F.silu(acc)
    elif epilogue == EPI_GATE_RESID:
        res = acc * _rows(gate, rows_per_gate)[:M] + resid.float()
    elif epilogue == EPI_RESID:
        res = acc + resid.float()
    elif epilogue == EPI_SWIGLU:
        h = _bf(acc).float()
        res = F.silu(h[:, :N // 2]) * h[:, N // 2:]
    elif epilogue == EPI_QKNORM:
        wq, wk, cos, sin, tokens = qknorm
        d = N // 3
        aux.copy_(qknorm_rope_fwd(_bf(acc), wq, wk, (cos, sin), d, tokens))
        res = acc
    else:
        res = acc
    if out is None:
        rows = out_rows if out_rows is not None else M
        out = torch.zeros((rows, res.shape[1]), device=A.device, dtype=out_dtype)
    if remap is not None:
        r, br, off = remap
        m = torch.arange(M, device=A.device)
        out[(m // r) * br + m % r + off] = res.to(out.dtype)
    elif accumulate:
        out += res.to(out.dtype)
    else:
        out.copy_(res)
    return out


# -------------------------------------------------------------------- attention
def qk_logit_bound(wq_x, wk_x, wq_c, wk_c, scale):
    m = max(float(wq_x.abs().max() * wk_x.abs().max()), float(wq_x.abs().max() * wk_c.abs().max()),
            float(wq_c.abs().max() * wk_x.abs().max()), float(wq_c.abs().max() * wk_c.abs().max()))
    return torch.full((1,), 64 * scale * m, device=wq_x.device, dtype=F32)


def _joint(tx, tc, B, H, N, M):
    a = tx.float().reshape(B, N, H, 64)
    if M:
        a = torch.cat([a, tc.float().reshape(B, M, H, 64)], 1)
    return a.permute(0, 2, 1, 3)


def _split(t, B, H, N, M):
    t = t.permute(0, 2, 1, 3).reshape(B, N + M, H * 64)
    return t[:, :N].reshape(B * N, H * 64), (t[:, N:].reshape(B * M, H * 64) if M else None)


def attn_fwd(q, k, v, B, H, N, M, scale, logit_bound=None):
    Q, K, V = (_joint(t[0], t[1], B, H, N, M) for t in (q, k, v))
    s = (Q @ K.transpose(-1, -2)) * scale
    o = s.softmax(-1) @ V
    o_x, o_c = _split(o, B, H, N, M)
    return _bf(o_x), (_bf(o_c) if M else None), torch.logsumexp(s, -1)


def attn_bwd(q, k, v, o, lse, d_o, dq, dk, dv, B, H, N, M, scale):
    with torch.enable_grad():
        Q, K, V = (_joint(t[0], t[1], B, H, N, M).detach().requires_grad_(True) for t in (q, k, v))
        out = ((Q @ K.transpose(-1, -2)) * scale).softmax(-1) @ V
        g = torch.autograd.grad(out, (Q, K, V), _joint(d_o[0], d_o[1], B, H, N, M))
    acc = None
    for grad, dst in zip(g, (dq, dk, dv)):
        gx, gc = _split(grad, B, H, N, M)
        if dst is None:         # dq stays in the fp32 accumulator (joint layout, image rows first)
            parts = [gx.float().view(B, N, H * 64)] + ([gc.float().view(B, M, H * 64)] if M else [])
            acc = torch.cat(parts, 1).contiguous()
            continue
        dst[0].copy_(gx)
        if M:
            dst[1].copy_(gc)
    return acc


def swiglu_bwd_fusable(rows, hidden):
    return rows > 128 and rows % 128 == 0 and hidden % 256 == 0


def gemm_swiglu_bwd(dy, w3b, h12, db12):
    return swiglu_bwd(gemm(dy, w3b, b_major=1), h12, db12)


# ------------------------------------------------------------------- row kernels
def _ln_mod(x, shift, scale, rpb):
    one_plus = (1 + scale.detach().to(BF16)).float() + (scale - scale.detach())   # bf16-rounded value, unit grad
    return F.layer_norm(x, (x.shape[-1],), eps=LN_EPS) * one_plus.repeat_interleave(rpb, 0) + \
        shift.repeat_interleave(rpb, 0)


def ln_modulate_fwd(x, shift, scale, rows_per_batch, save_stats=True):
    xf = x.float()
    y = _bf(_ln_mod(xf, shift.float(), scale.float(), rows_per_batch))
    mean = xf.mean(-1)
    rstd = torch.rsqrt(xf.var(-1, unbiased=False) + LN_EPS)
    return y, (mean if save_stats else None), (rstd if save_stats else None)


def ln_modulate_bwd(dy, x, mean, rstd, scale, dres, dshift, dscale, rows_per_batch):
    B = scale.shape[0]
    with torch.enable_grad():
        xf = x.float().detach().requires_grad_(True)
        sc = scale.float().detach().clone().requires_grad_(True)
        sh = torch.zeros_like(sc).requires_grad_(True)
        y = _ln_mod(xf, sh, sc, rows_per_batch)
        gx, gsh, gsc = torch.autograd.grad(y, (xf, sh, sc), dy.float())
    if dres is not None:
        gx = gx + dres.float()
    dshift.copy_(gsh.view(B, -1))
    dscale.copy_(gsc.view(B, -1))
    return _bf(gx)


def gate_residual_fwd(a, gate, resid, rows_per_batch):
    return _bf(a.float() * _rows(gate, rows_per_batch) + resid.float())


def gate_residual_ln_fwd(a, gate, resid, shift, scale, rows_per_batch):
    xo = gate_residual_fwd(a, gate, resid, rows_per_batch)
    y, mean, rstd = ln_modulate_fwd(xo, shift, scale, rows_per_batch)
    return xo, y, mean, rstd


def gate_bwd(dout, a, gate, dgate, dab, rows_per_batch):
    B, d = gate.shape
    da = dout.float() * _rows(gate, rows_per_batch)
    dgate.copy_((dout.float() * a.float()).view(B, rows_per_batch, d).sum(1))
    if dab is not None:
        dab.copy_(da.view(B, rows_per_batch, d).sum(1))
    return _bf(da)


def text_norm_fwd(c, w1, w2, s1, s2, split):
    Bn, M, dt = c.shape
    cf = c.float()
    o1 = s1.float() * F.rms_norm(cf[:, :split], (dt,), w1.float(), RMS_EPS)
    o2 = s2.float() * F.rms_norm(cf[:, split:], (dt,), w2.float(), RMS_EPS) if M > split else None
    rstd = torch.rsqrt(cf.pow(2).mean(-1) + RMS_EPS).reshape(Bn * M)
    return _bf(o1.reshape(-1, dt)), (_bf(o2.reshape(-1, dt)) if o2 is not None else None), rstd


def text_norm_bwd(dn, c, rstd, w, sigma, dw, dsigma, tok0, ntok):
    Bn, M, dt = c.shape
    with torch.enable_grad():
        wf = w.detach().float().requires_grad_(True)
        sf = sigma.detach().float().requires_grad_(True)
        out = sf * F.rms_norm(c[:, tok0:tok0 + ntok].float(), (dt,), wf, RMS_EPS)
        gw, gs = torch.autograd.grad((out.reshape(-1, dt) * dn.float()).sum(), (wf, sf))
    dw += gw
    dsigma += gs


def swiglu_fwd(h12):
    hid = h12.shape[1] // 2
    h = h12.float()
    return _bf(F.silu(h[:, :hid]) * h[:, hid:])


def swiglu_bwd(da, h12, db12):
    hid = h12.shape[1] // 2
    with torch.enable_grad():
        h = h12.float().detach().requires_grad_(True)
        (g,) = torch.autograd.grad(F.silu(h[:, :hid]) * h[:, hid:], h, da.float())
    if db12 is not None:
        db12 += g.sum(0)
    return _bf(g)


def _temb(t, ts, denom):
    emb = (t.float() * ts.float())[:, None] / denom.float()[None]
    return torch.cat((emb[:, ::2].sin(), emb[:, 1::2].cos()), 1)


def timestep_embed_fwd(t, time_scale, denom):
    return _bf(_temb(t, time_scale, denom))


def timestep_embed_bwd(de, t, time_scale, denom, dscale):
    with torch.enable_grad():
        ts = time_scale.detach().float().requires_grad_(True)
        (g,) = torch.autograd.grad((_temb(t, ts, denom) * de.float()).sum(), ts)
    dscale += g


# ------------------------------------------------------------------- elementwise
def patchify(img, p):
    Bn, Cc, H, W = img.shape
    return _bf(F.unfold(img.float(), kernel_size=p, stride=p).transpose(1, 2).reshape(-1, Cc * p * p))


def unpatchify(tok, Bn, Cc, H, W, p, dtype=BF16):
    cols = tok.float().reshape(Bn, (H // p) * (W // p), Cc * p * p).transpose(1, 2)
    return F.fold(cols, (H, W), kernel_size=p, stride=p).to(dtype)


def rf_noise(x0, eps, t):
    tt = t.float()[:, None, None, None]
    return (1 - tt) * x0.float() + tt * eps.float()


def rf_loss_fwd(v, eps, x0):
    diff = v.float() - (eps.float() - x0.float())
    return diff.pow(2).mean(), diff


def rf_loss_bwd(diff, upstream, dtype):
    return (upstream.float() * 2.0 * diff / diff.numel()).to(dtype)


def cfg_euler_step(x, v, cfg_scale, dt):
    B = x.shape[0]
    x -= ((1 + cfg_scale) * v[:B].float() - cfg_scale * v[B:].float()) * dt
    return x


def colsum(x, out=None):
    if out is None:
        out = torch.zeros(x.shape[1], device=x.device, dtype=F32)
    out += x.float().sum(0)
    return out


def fold_rows(x, out):
    out += x.sum(0)
    return out


def cast_bf16(x, out=None):
    if out is None:
        return x.to(BF16)
    out.copy_(x)
    return out


ALL = [n for n, f in list(globals().items()) if callable(f) and not n.startswith("_") and n not in ("F",)]

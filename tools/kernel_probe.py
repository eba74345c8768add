"""GPU diagnostic for the attention and memory-bound kernels: each op is compared
with a plain torch fp32 restatement on the device.  python tools/kernel_probe.py <group>
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib, ops  # noqa: E402

dev = "cuda"
BF = torch.bfloat16


def rel(got, ref, tag, tol=2e-2):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-20
    err = (got - ref).abs().max().item() / scale
    bad = not (err < tol)
    print(f"    {tag:28s} rel={err:.3e} ref_max={scale:.3e} {'FAIL' if bad else 'ok'}")
    if bad:
        d = (got - ref).abs()
        idx = torch.nonzero(d > tol * scale)
        print("      n_bad", idx.shape[0], "of", d.numel(), "first", idx[:4].tolist(),
              "nan", int(torch.isnan(got).sum()))
    return not bad


def ref_attention(q, k, v, scale):
    # q,k,v: [B,H,T,64] fp32
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return p @ v, torch.logsumexp(s, -1)


def _unit_rms_heads(t, d, H):
    # q and k as they leave QK-RMSNorm (weight 1): every 64-wide head has unit RMS
    qk = t[:, :2 * d].float().view(t.shape[0], 2 * H, 64)
    qk = qk * torch.rsqrt(qk.pow(2).mean(-1, keepdim=True))
    t[:, :2 * d] = qk.view(t.shape[0], 2 * d).bfloat16()
    return t


def attn_case(B, H, N, M, bwd=True, seed=0, bounded=False, unusable_bound=False):
    torch.manual_seed(seed)
    d = H * 64
    print(f"[attn] B={B} H={H} N={N} M={M} bounded={bounded} unusable_bound={unusable_bound}")
    qkv_x = torch.randn(B * N, 3 * d, device=dev).bfloat16()
    qkv_c = torch.randn(B * M, 3 * d, device=dev).bfloat16() if M else None
    bound = None
    if bounded:
        qkv_x = _unit_rms_heads(qkv_x, d, H)
        qkv_c = _unit_rms_heads(qkv_c, d, H) if M else None
        one = torch.ones(64, device=dev)
        bound = ops.qk_logit_bound(one, one, one, one, 0.125)
        assert abs(float(bound) - 8.16) < 1e-4
    if unusable_bound:   # a bound above 24: attn_fwd2 returns at once, the online-softmax kernel's item loop runs
        bound = torch.full((1,), 30.0, device=dev)
    qs = (qkv_x[:, :d], qkv_c[:, :d] if M else None)
    ks = (qkv_x[:, d:2 * d], qkv_c[:, d:2 * d] if M else None)
    vs = (qkv_x[:, 2 * d:], qkv_c[:, 2 * d:] if M else None)
    scale = 0.125
    o_x, o_c, lse = ops.attn_fwd(qs, ks, vs, B, H, N, M, scale, logit_bound=bound)
    torch.cuda.synchronize()

    def joint(tx, tc):
        a = tx.float().reshape(B, N, H, 64)
        if M:
            a = torch.cat([a, tc.float().reshape(B, M, H, 64)], 1)
        return a.permute(0, 2, 1, 3).contiguous()

    q, k, v = joint(*qs), joint(*ks), joint(*vs)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    o_ref, lse_ref = ref_attention(q, k, v, scale)
    o_got = joint(o_x, o_c)
    ok = rel(o_got, o_ref, "o")
    ok &= rel(lse, lse_ref, "lse", 1e-3)
    if not bwd:
        return ok
    do_x = torch.randn(B * N, d, device=dev).bfloat16()
    do_c = torch.randn(B * M, d, device=dev).bfloat16() if M else None
    dqkv_x = torch.zeros_like(qkv_x)
    dqkv_c = torch.zeros_like(qkv_c) if M else None
    dq = (dqkv_x[:, :d], dqkv_c[:, :d] if M else None)
    dk = (dqkv_x[:, d:2 * d], dqkv_c[:, d:2 * d] if M else None)
    dv = (dqkv_x[:, 2 * d:], dqkv_c[:, 2 * d:] if M else None)
    ops.attn_bwd(qs, ks, vs, (o_x, o_c), lse, (do_x, do_c), dq, dk, dv, B, H, N, M, scale)
    torch.cuda.synchronize()
    o_ref.backward(joint(do_x, do_c))
    ok &= rel(joint(*dq), q.grad, "dq")
    ok &= rel(joint(*dk), k.grad, "dk")
    ok &= rel(joint(*dv), v.grad, "dv")
    return ok


def group_attn():
    ok = True
    ok &= attn_case(1, 1, 128, 0)
    ok &= attn_case(1, 2, 256, 128)
    ok &= attn_case(2, 4, 256, 154)
    ok &= attn_case(2, 3, 240, 77)      # 24x40 latent is 12x20 tokens; ragged everywhere
    ok &= attn_case(1, 2, 1024, 256)
    ok &= attn_case(2, 4, 256, 154, bounded=True)      # single-pass softmax against the QK-norm bound
    ok &= attn_case(2, 3, 240, 77, bounded=True)
    ok &= attn_case(2, 4, 1024, 154, bounded=True)     # 10 key tiles, 26-row last query tile (idle softmax warps)
    ok &= attn_case(8, 12, 256, 154, bwd=False, unusable_bound=True)   # 384 items on 296 looping CTAs
    ok &= attn_case(2, 3, 240, 77, bwd=False, unusable_bound=True)
    return ok


def group_attn_perf():
    for (B, H, N, M) in [(64, 12, 256, 154), (16, 24, 1024, 154)]:
        d = H * 64
        qkv_x = torch.randn(B * N, 3 * d, device=dev).bfloat16()
        qkv_c = torch.randn(B * M, 3 * d, device=dev).bfloat16()
        qkv_x, qkv_c = _unit_rms_heads(qkv_x, d, H), _unit_rms_heads(qkv_c, d, H)
        one = torch.ones(64, device=dev)
        bound = ops.qk_logit_bound(one, one, one, one, 0.125)
        qs, ks, vs = ((qkv_x[:, i * d:(i + 1) * d], qkv_c[:, i * d:(i + 1) * d]) for i in range(3))
        o_x, o_c, lse = ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125)
        do_x, do_c = torch.randn_like(o_x), torch.randn_like(o_c)
        dx, dc = torch.empty_like(qkv_x), torch.empty_like(qkv_c)
        dq, dk, dv = ((dx[:, i * d:(i + 1) * d], dc[:, i * d:(i + 1) * d]) for i in range(3))
        T = N + M
        flops = 4.0 * B * H * T * T * 64

        def timeit(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        ms = timeit(lambda: ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125))
        print(f"[attn perf] B={B} H={H} T={T} fwd (online softmax) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TFLOP/s")
        ms = timeit(lambda: ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound))
        print(f"[attn perf] B={B} H={H} T={T} fwd (bounded, 1 pass) {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TFLOP/s")
        ms = timeit(lambda: ops.attn_bwd(qs, ks, vs, (o_x, o_c), lse, (do_x, do_c), dq, dk, dv, B, H,
                                         N, M, 0.125))
        print(f"[attn perf] B={B} H={H} T={T} bwd {ms * 1e3:8.1f} us {2.5 * flops / ms / 1e9:7.1f} TFLOP/s")
        # library reference: torch SDPA on the joint sequence
        q = torch.randn(B, H, T, 64, device=dev, dtype=BF)
        ms = timeit(lambda: F.scaled_dot_product_attention(q, q, q, scale=0.125))
        print(f"[attn perf] B={B} H={H} T={T} torch SDPA fwd {ms * 1e3:8.1f} us {flops / ms / 1e9:7.1f} TFLOP/s")
    return True


def group_rowwise():
    ok = True
    torch.manual_seed(1)
    for (B, rpb, d) in [(4, 154, 768), (2, 256, 256), (3, 77, 1216), (2, 64, 1536)]:
        print(f"[ln_modulate] B={B} rows/batch={rpb} d={d}")
        R = B * rpb
        x = torch.randn(R, d, device=dev).bfloat16()
        mod = (torch.randn(B, 3 * d, device=dev) * 0.5).bfloat16()
        shift, scale = mod[:, :d], mod[:, d:2 * d]
        y, mean, rstd = ops.ln_modulate_fwd(x, shift, scale, rpb)
        xf = x.float().requires_grad_(True)
        sh = shift.float().clone().requires_grad_(True)
        sc = scale.float().clone().requires_grad_(True)
        one_plus = (1 + sc.detach().bfloat16()).float() + (sc - sc.detach())  # bf16-rounded value, unit grad
        yr = F.layer_norm(xf, (d,)) * one_plus.repeat_interleave(rpb, 0) + sh.repeat_interleave(rpb, 0)
        ok &= rel(y, yr, "fwd y", 1e-2)
        dy = torch.randn(R, d, device=dev).bfloat16()
        dres = torch.randn(R, d, device=dev).bfloat16()
        dmod = torch.zeros(B, 2 * d, device=dev)
        dx = ops.ln_modulate_bwd(dy, x, mean, rstd, scale, dres, dmod[:, :d], dmod[:, d:], rpb)
        yr.backward(dy.float())
        ok &= rel(dx, xf.grad + dres.float(), "bwd dx", 1e-2)
        ok &= rel(dmod[:, :d], sh.grad, "bwd dshift", 5e-3)
        ok &= rel(dmod[:, d:], sc.grad, "bwd dscale", 5e-3)

        print(f"[gate_bwd] B={B} rows/batch={rpb} d={d}")
        a = torch.randn(R, d, device=dev).bfloat16()
        g = torch.randn(B, d, device=dev).bfloat16()
        dout = torch.randn(R, d, device=dev).bfloat16()
        dg = torch.zeros(B, d, device=dev)
        dab = torch.zeros(B, d, device=dev)
        da = ops.gate_bwd(dout, a, g, dg, dab, rpb)
        da_ref = dout.float() * g.float().repeat_interleave(rpb, 0)
        ok &= rel(da, da_ref, "da", 1e-2)
        ok &= rel(dg, (dout.float() * a.float()).view(B, rpb, d).sum(1), "dgate", 5e-3)
        ok &= rel(dab, da_ref.view(B, rpb, d).sum(1), "dab", 5e-3)

    print("[text_norm]")
    B, M, dt, split = 3, 154, 2304, 77
    c = (torch.randn(B, M, dt, device=dev) * 5).bfloat16()
    w1 = torch.rand(dt, device=dev) + 0.5
    w2 = torch.rand(dt, device=dev) + 0.5
    s1 = torch.tensor([0.01], device=dev)
    s2 = torch.tensor([0.02], device=dev)
    o1, o2, rstd = ops.text_norm_fwd(c, w1, w2, s1, s2, split)
    w1r, w2r = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    s1r, s2r = s1.clone().requires_grad_(True), s2.clone().requires_grad_(True)
    r1 = s1r * F.rms_norm(c[:, :split].float(), (dt,), w1r, None)
    r2 = s2r * F.rms_norm(c[:, split:].float(), (dt,), w2r, None)
    ok &= rel(o1, r1.reshape(-1, dt), "fwd half1", 1e-2)
    ok &= rel(o2, r2.reshape(-1, dt), "fwd half2", 1e-2)
    dn1 = torch.randn(B * split, dt, device=dev).bfloat16()
    dn2 = torch.randn(B * (M - split), dt, device=dev).bfloat16()
    dw1, dw2 = torch.zeros(dt, device=dev), torch.zeros(dt, device=dev)
    ds1, ds2 = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    ops.text_norm_bwd(dn1, c, rstd, w1, s1, dw1, ds1, 0, split)
    ops.text_norm_bwd(dn2, c, rstd, w2, s2, dw2, ds2, split, M - split)
    (r1.reshape(-1, dt) * dn1.float()).sum().backward()
    (r2.reshape(-1, dt) * dn2.float()).sum().backward()
    ok &= rel(dw1, w1r.grad, "dw1", 5e-3)
    ok &= rel(dw2, w2r.grad, "dw2", 5e-3)
    ok &= rel(ds1, s1r.grad, "dsigma1", 5e-3)
    ok &= rel(ds2, s2r.grad, "dsigma2", 5e-3)
    return ok


def rope_tables(h, w, freqs):
    # rotary_embedding.py:269-288 get_axial_freqs + :72 cos/sin, per interleaved pair
    fh = torch.arange(h, device=dev).float()[:, None] * freqs[None]      # [h,16]
    fw = torch.arange(w, device=dev).float()[:, None] * freqs[None]      # [w,16]
    ang = torch.cat([fh[:, None, :].expand(h, w, 16), fw[None, :, :].expand(h, w, 16)], -1)
    ang = ang.reshape(h * w, 32)
    return ang.cos().contiguous(), ang.sin().contiguous()


def ref_qknorm_rope(x, w, H, rope, tokens):
    # x [R, d] fp32 -> per-head RMSNorm (bf16-rounded) then interleaved-pair rotation
    R, d = x.shape
    xh = x.view(R, H, 64)
    n = F.rms_norm(xh, (64,), w, None)
    n = n + (n.bfloat16().float() - n).detach()
    if rope is None:
        return n.reshape(R, d)
    cos, sin = rope
    tok = torch.arange(R, device=dev) % tokens
    c = cos[tok].repeat_interleave(2, -1)[:, None, :]
    s = sin[tok].repeat_interleave(2, -1)[:, None, :]
    x1, x2 = n[..., 0::2], n[..., 1::2]
    rot = torch.stack((-x2, x1), -1).reshape(R, H, 64)
    return (n * c + rot * s).reshape(R, d)


def group_elem():
    ok = True
    torch.manual_seed(2)
    for (B, h, w, H, img) in [(2, 16, 16, 4, True), (2, 12, 20, 12, True), (3, 154, 1, 4, False)]:
        tokens, d = h * w, H * 64
        R = B * tokens
        print(f"[qknorm_rope] B={B} tokens={tokens} H={H} rope={img}")
        qkv = torch.randn(R, 3 * d, device=dev).bfloat16()
        wq, wk = torch.rand(64, device=dev) + 0.5, torch.rand(64, device=dev) + 0.5
        freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2, device=dev).float() / 32))
        rope = rope_tables(h, w, freqs) if img else None
        out = ops.qknorm_rope_fwd(qkv, wq, wk, rope, d, tokens)
        qf = qkv[:, :d].float().requires_grad_(True)
        kf = qkv[:, d:2 * d].float().requires_grad_(True)
        wqr, wkr = wq.clone().requires_grad_(True), wk.clone().requires_grad_(True)
        qr = ref_qknorm_rope(qf, wqr, H, rope, tokens)
        kr = ref_qknorm_rope(kf, wkr, H, rope, tokens)
        ok &= rel(out[:, :d], qr, "fwd q", 1e-2)
        ok &= rel(out[:, d:], kr, "fwd k", 1e-2)
        dqk = torch.randn(R, 2 * d, device=dev).bfloat16()
        dqkv = torch.zeros(R, 3 * d, device=dev, dtype=BF)
        dwq, dwk = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
        ops.qknorm_rope_bwd(dqk, qkv, wq, wk, rope, dqkv, dwq, dwk, d, tokens)
        (qr * dqk[:, :d].float()).sum().backward()
        (kr * dqk[:, d:].float()).sum().backward()
        ok &= rel(dqkv[:, :d], qf.grad, "bwd dq", 1e-2)
        ok &= rel(dqkv[:, d:2 * d], kf.grad, "bwd dk", 1e-2)
        ok &= rel(dwq, wqr.grad, "bwd dwq", 5e-3)
        ok &= rel(dwk, wkr.grad, "bwd dwk", 5e-3)
        # q half of the gradient read from an fp32 accumulator in the attention backward's joint layout
        # [B, T, d] (this stream's rows at an offset inside every sample): must equal the bf16-copy path
        pad = 7
        acc = torch.randn(B, tokens + 2 * pad, d, device=dev)
        acc[:, pad:pad + tokens] = dqk[:, :d].float().view(B, tokens, d)
        dqkv2 = torch.zeros(R, 3 * d, device=dev, dtype=BF)
        dwq2, dwk2 = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
        dqk_k_only = dqk.clone()
        dqk_k_only[:, :d] = float("nan")       # the q half of dqk must not be read
        ops.qknorm_rope_bwd(dqk_k_only, qkv, wq, wk, rope, dqkv2, dwq2, dwk2, d, tokens, dq_acc=acc, acc_off=pad)
        same = bool(torch.equal(dqkv2[:, :2 * d], dqkv[:, :2 * d]))
        print(f"    bwd from fp32 accumulator     identical to the bf16-copy path: {same}")
        ok &= same
        ok &= rel(dwq2, dwq, "bwd dwq (acc path)", 1e-5)

    print("[swiglu]")
    R, hid = 308, 1024
    h12 = torch.randn(R, 2 * hid, device=dev).bfloat16()
    a = ops.swiglu_fwd(h12)
    hf = h12.float().requires_grad_(True)
    ar = F.silu(hf[:, :hid]) * hf[:, hid:]
    ok &= rel(a, ar, "fwd", 1e-2)
    da = torch.randn(R, hid, device=dev).bfloat16()
    db = torch.zeros(2 * hid, device=dev)
    dh = ops.swiglu_bwd(da, h12, db)
    ar.backward(da.float())
    ok &= rel(dh, hf.grad, "bwd dh12", 1e-2)
    ok &= rel(db, hf.grad.sum(0), "bwd db12", 5e-3)

    print("[timestep_embed]")
    B, d = 5, 256
    t = torch.rand(B, device=dev)
    ts = torch.tensor([1000.0], device=dev, requires_grad=True)
    denom = (torch.tensor(10000.0) ** ((2 * torch.arange(d)) / d)).to(dev)
    e = ops.timestep_embed_fwd(t, ts.detach(), denom)
    emb = (t * ts)[:, None] / denom[None]
    er = torch.cat((emb[:, ::2].sin(), emb[:, 1::2].cos()), 1)
    ok &= rel(e, er, "fwd", 1e-2)
    de = torch.randn(B, d, device=dev).bfloat16()
    dts = torch.zeros(1, device=dev)
    ops.timestep_embed_bwd(de, t, ts.detach(), denom, dts)
    (er * de.float()).sum().backward()
    ok &= rel(dts, ts.grad, "bwd dscale", 5e-3)

    print("[patchify / unpatchify]")
    B, Cc, H, W, p = 2, 16, 24, 40, 2
    img = torch.randn(B, Cc, H, W, device=dev)
    tok = ops.patchify(img, p)
    ref_tok = F.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(-1, Cc * p * p)
    ok &= rel(tok, ref_tok, "patchify (vs unfold)", 1e-2)
    back = ops.unpatchify(tok, B, Cc, H, W, p, torch.float32)
    ok &= rel(back, img, "unpatchify round trip", 1e-2)

    print("[rf noise / loss / cfg euler]")
    x0 = torch.randn(4, 16, 32, 32, device=dev).bfloat16()
    eps = torch.randn_like(x0)
    t = torch.rand(4, device=dev)
    xt = ops.rf_noise(x0, eps, t)
    tt = t[:, None, None, None]
    ok &= rel(xt, (1 - tt) * x0.float() + tt * eps.float(), "rf_noise", 1e-5)
    v = torch.randn(4, 16, 32, 32, device=dev).bfloat16()
    loss, diff = ops.rf_loss_fwd(v, eps, x0)
    vr = v.float().requires_grad_(True)
    lr = F.mse_loss(vr, eps.float() - x0.float())
    ok &= rel(loss, lr, "rf_loss", 1e-5)
    dv = ops.rf_loss_bwd(diff, torch.tensor(3.0, device=dev), torch.bfloat16)
    (lr * 3.0).backward()
    ok &= rel(dv, vr.grad, "rf_loss dv", 1e-2)
    x = torch.randn(2, 16, 32, 32, device=dev)
    xr = x - ((1 + 5.0) * v[:2].float() - 5.0 * v[2:].float()) * 0.02
    ops.cfg_euler_step(x, v, 5.0, 0.02)
    ok &= rel(x, xr, "cfg_euler", 1e-5)

    print("[colsum / fold / cast]")
    m = torch.randn(1000, 776, device=dev).bfloat16()
    ok &= rel(ops.colsum(m), m.float().sum(0), "colsum", 1e-3)
    f = torch.randn(7, 300, device=dev)
    o = torch.ones(300, device=dev)
    ok &= rel(ops.fold_rows(f, o), f.sum(0) + 1, "fold_rows", 1e-5)
    z = torch.randn(12345, device=dev)
    ok &= rel(ops.cast_bf16(z), z.bfloat16(), "cast", 1e-6)
    return ok


if __name__ == "__main__":
    g = sys.argv[1]
    _lib.check(_lib.lib().mmdit_device_check(), "device_check")
    t0 = time.time()
    fn = {"attn": group_attn, "attn_perf": group_attn_perf, "rowwise": group_rowwise,
          "elem": group_elem}[g]
    ok = fn()
    print(f"GROUP {g}: {'ALL PASS' if ok else 'SOME FAIL'} ({time.time() - t0:.1f}s)")
    sys.exit(0 if ok else 1)

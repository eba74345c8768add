"""GB/s table of the LayerNorm-modulate / gate / QK-norm row kernels at the shapes the train step runs them
at (CUDA-graph replay over rotating buffer sets larger than the L2), both kernel generations side by side;
a bit-for-bit check of the fused gate+residual+LN forward against the two kernels it replaces; and every
second-generation kernel against its first-generation counterpart on the same inputs.
python tools/row_probe.py [check|perf|all]   (exit code 1 on any mismatch)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops  # noqa: E402

dev = "cuda"
BF, F32 = torch.bfloat16, torch.float32
HBM_PEAK = 6549.8   # GB/s, MEASURED_PEAKS.json copy bandwidth
ok_all = True


def rel(got, ref, tag, tol):
    global ok_all
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-20
    err = (got - ref).abs().max().item() / scale
    bad = not (err < tol)
    ok_all &= not bad
    print(f"    {tag:34s} rel={err:.3e} {'FAIL' if bad else 'ok'}")


def same(got, ref, tag):
    """bit-identical to the first-generation kernel (same math, same summation order per row)"""
    n = int((got.view(torch.int16) != ref.view(torch.int16)).sum()) if got.dtype == BF else int((got != ref).sum())
    print(f"    {tag:34s} differing elements: {n} of {got.numel()}")
    return n


def close_bf16(got, ref, tag, exact):
    """second generation vs first: bit-identical where the operation order is the same (forward kernels), else
    at most one bf16 ulp (plus the fp32 noise of a column sum) apart: the strips of the two generations are cut differently"""
    global ok_all
    g, r = got.float(), ref.float()
    diff = (g - r).abs()
    n = int((diff > 0).sum())
    # column sums cancel: an entry near zero carries the fp32 noise of the whole sum
    noise = 2e-5 * float(r.abs().max()) + 1e-12
    ulp = torch.maximum(g.abs(), r.abs()) * 2.0 ** -7 + noise
    worst = float((diff / ulp).max()) if n else 0.0
    if exact:
        bad = n != 0
    elif got.dtype == BF:
        bad = worst > 1.001
    else:
        bad = float(diff.max()) > noise
    ok_all &= not bad
    print(f"    {tag:34s} differing: {n} of {got.numel()} (worst {worst:.2f} bf16 ulp) {'FAIL' if bad else 'ok'}")


def both_generations(fn):
    ops.set_row_kernel_generation(1)
    a = fn()
    ops.set_row_kernel_generation(2)
    b = fn()
    return a, b


def check_generations(Bn, T, d):
    """every second-generation row kernel against the first-generation one on the same inputs"""
    print(f"[gen 2 vs gen 1] B={Bn} rows/sample={T} d={d}")
    t = make(Bn, T, d, seed=3)
    a, resid, gate, shift, scale, dy, dres = (t[k] for k in ("a", "resid", "gate", "shift", "scale", "dy", "dres"))
    (y1, m1, r1), (y2, m2, r2) = both_generations(lambda: ops.ln_modulate_fwd(a, shift, scale, T))
    close_bf16(y2, y1, "ln_modulate_fwd y", True)
    close_bf16(m2, m1, "ln_modulate_fwd mean", True)
    close_bf16(r2, r1, "ln_modulate_fwd rstd", True)
    if d <= 1024:
        f1, f2 = both_generations(lambda: ops.gate_residual_ln_fwd(a, gate, resid, shift, scale, T))
        for k, nm in enumerate(("x'", "y", "mean", "rstd")):
            close_bf16(f2[k], f1[k], f"gate_residual_ln_fwd {nm}", True)

    def lnb():
        dmod = torch.zeros(2, Bn, d, device=dev, dtype=BF)
        dx = ops.ln_modulate_bwd(dy, a, m1, r1, scale, dres, dmod[0], dmod[1], T)
        dmod32 = torch.zeros(2, Bn, d, device=dev)
        dx0 = ops.ln_modulate_bwd(dy, a, m1, r1, scale, None, dmod32[0], dmod32[1], T)
        return dx, dmod, dx0, dmod32
    b1, b2 = both_generations(lnb)
    for k, nm in enumerate(("dx (+dres)", "dshift/dscale bf16", "dx (no dres)", "dshift/dscale fp32")):
        close_bf16(b2[k], b1[k], f"ln_modulate_bwd {nm}", False)

    def gb():
        dgate = torch.zeros(Bn, d, device=dev, dtype=BF)
        dab = torch.zeros(Bn, d, device=dev)
        da = ops.gate_bwd(dy, a, gate, dgate, dab, T)
        return da, dgate, dab
    g1, g2 = both_generations(gb)
    for k, nm in enumerate(("da", "dgate", "dab")):
        close_bf16(g2[k], g1[k], f"gate_bwd {nm}", nm == "da")

    def sb():
        gg = torch.Generator(device=dev).manual_seed(5)
        h12 = torch.randn(Bn * T, 4 * d, device=dev, generator=gg).bfloat16()
        dact = torch.randn(Bn * T, 2 * d, device=dev, generator=gg).bfloat16()
        db = torch.zeros(4 * d, device=dev)
        return ops.swiglu_bwd(dact, h12, db), db
    s1, s2 = both_generations(sb)
    close_bf16(s2[0], s1[0], "swiglu_bwd dh12", True)
    close_bf16(s2[1], s1[1], "swiglu_bwd db12", False)   # atomics: fp32 order varies

    if d % 64 == 0:
        g = torch.Generator(device=dev).manual_seed(11)
        R = Bn * T
        qkv = torch.randn(R, 3 * d, device=dev, generator=g).bfloat16()
        dqk = torch.randn(R, 2 * d, device=dev, generator=g).bfloat16()
        wq, wk = torch.rand(64, device=dev, generator=g) + 0.5, torch.rand(64, device=dev, generator=g) + 0.5
        ang = torch.rand(T, 32, device=dev, generator=g) * 6.0
        for rope in ((torch.cos(ang).contiguous(), torch.sin(ang).contiguous()), None):
            tag = "rope" if rope else "text"
            o1, o2 = both_generations(lambda: ops.qknorm_rope_fwd(qkv, wq, wk, rope, d, T))
            close_bf16(o2, o1, f"qknorm_rope_fwd ({tag})", True)

            def qb(acc):
                dqkv = torch.zeros(R, 3 * d, device=dev, dtype=BF)
                dw = torch.zeros(2, 64, device=dev)
                if acc:
                    pad = 5
                    dq_acc = torch.zeros(Bn, T + 2 * pad, d, device=dev)
                    dq_acc[:, pad:pad + T] = dqk[:, :d].float().view(Bn, T, d)
                    ops.qknorm_rope_bwd(dqk, qkv, wq, wk, rope, dqkv, dw[0], dw[1], d, T, dq_acc=dq_acc, acc_off=pad)
                else:
                    ops.qknorm_rope_bwd(dqk, qkv, wq, wk, rope, dqkv, dw[0], dw[1], d, T)
                return dqkv[:, :2 * d], dw
            for acc in (False, True):
                q1, q2 = both_generations(lambda: qb(acc))
                close_bf16(q2[0], q1[0], f"qknorm_rope_bwd ({tag}{', fp32 dq' if acc else ''}) dqk", False)
                close_bf16(q2[1], q1[1], f"qknorm_rope_bwd ({tag}{', fp32 dq' if acc else ''}) dw", False)


def make(Bn, T, d, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    R = Bn * T
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    t = dict(a=r(R, d).bfloat16(), resid=r(R, d).bfloat16(), dy=r(R, d).bfloat16(), dres=r(R, d).bfloat16(),
             mod=(0.3 * r(Bn, 4 * d)).bfloat16())
    t["shift"], t["scale"], t["gate"] = t["mod"][:, :d], t["mod"][:, d:2 * d], t["mod"][:, 2 * d:3 * d]
    return t


def check(Bn, T, d):
    global ok_all
    print(f"[rows] B={Bn} rows/sample={T} d={d}")
    t = make(Bn, T, d)
    a, resid, gate, shift, scale = (t[k] for k in ("a", "resid", "gate", "shift", "scale"))
    xo, y, mean, rstd = ops.gate_residual_ln_fwd(a, gate, resid, shift, scale, T)
    xo1 = ops.gate_residual_fwd(a, gate, resid, T)
    y1, mean1, rstd1 = ops.ln_modulate_fwd(xo1, shift, scale, T)
    n = same(xo, xo1, "fused fwd x'") + same(y, y1, "fused fwd y") + same(mean, mean1, "mean") + same(rstd, rstd1, "rstd")
    ok_all &= n == 0
    xf = (a.float().view(Bn, T, d) * gate.float()[:, None] + resid.float().view(Bn, T, d)).bfloat16().float()
    mu, var = xf.mean(-1, keepdim=True), xf.var(-1, unbiased=False, keepdim=True)
    yf = (xf - mu) * torch.rsqrt(var + 1e-5) * (1 + scale).float()[:, None] + shift.float()[:, None]
    rel(y.view(Bn, T, d), yf, "fused fwd y vs fp32", 1e-2)


def timeit(fn, nbuf, iters=5):
    """fn(i) runs on buffer set i; sets rotate so that every call reads data that has left the L2.
    The calls are captured into one CUDA graph (the Python / ctypes launch cost of a 15 us kernel would
    otherwise be what is measured) and the graph is replayed."""
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(nbuf):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * nbuf) * 1e3   # us


def perf(name, Bn, T, d):
    R = Bn * T
    per = R * d * 2   # bytes of one bf16 [R, d] tensor
    nbuf = max(2, int(400e6 // (5 * per)) + 1)   # > 126 MB of L2 between two uses of the same set (row kernels)
    sets = [make(Bn, T, d, seed=i) for i in range(nbuf)]
    for s in sets:
        s["xo"], s["y"], s["mean"], s["rstd"] = ops.gate_residual_ln_fwd(s["a"], s["gate"], s["resid"], s["shift"], s["scale"], T)
        s["dmod"] = torch.empty(2, Bn, d, device=dev, dtype=BF)
        s["dgate"] = torch.empty(Bn, d, device=dev, dtype=BF)
        s["dab"] = torch.empty(Bn, d, device=dev, dtype=F32)
    rows = []

    def add(tag, nbytes, fn):
        us = []
        for gen in (1, 2):
            ops.set_row_kernel_generation(gen)
            us.append(timeit(fn, nbuf))
        rows.append((tag, us, [nbytes / u / 1e3 for u in us]))

    S = lambda i: sets[i]
    add("gate_residual_fwd", 3 * per, lambda i: ops.gate_residual_fwd(S(i)["a"], S(i)["gate"], S(i)["resid"], T))
    add("ln_modulate_fwd", 2 * per, lambda i: ops.ln_modulate_fwd(S(i)["xo"], S(i)["shift"], S(i)["scale"], T))
    add("gate_residual_ln_fwd", 4 * per, lambda i: ops.gate_residual_ln_fwd(
        S(i)["a"], S(i)["gate"], S(i)["resid"], S(i)["shift"], S(i)["scale"], T))
    add("ln_modulate_bwd (dres)", 4 * per, lambda i: ops.ln_modulate_bwd(
        S(i)["dy"], S(i)["xo"], S(i)["mean"], S(i)["rstd"], S(i)["scale"], S(i)["dres"], S(i)["dmod"][0], S(i)["dmod"][1], T))
    add("gate_bwd", 3 * per, lambda i: ops.gate_bwd(S(i)["dy"], S(i)["a"], S(i)["gate"], S(i)["dgate"], S(i)["dab"], T))
    # QK-RMSNorm + RoPE and SwiGLU backward at the same row count (image stream: RoPE on)
    H = d // 64
    side = int(round(T ** 0.5))
    rope = None
    if side * side == T:
        ang = torch.rand(T, 32, device=dev)
        rope = (torch.cos(ang).contiguous(), torch.sin(ang).contiguous())
    wq, wk = torch.rand(64, device=dev) + 0.5, torch.rand(64, device=dev) + 0.5
    for s in sets:
        g = torch.Generator(device=dev).manual_seed(7)
        s["qkv"] = torch.randn(R, 3 * d, device=dev, generator=g).bfloat16()
        s["dqk"] = torch.randn(R, 2 * d, device=dev, generator=g).bfloat16()
        s["dqkv"] = torch.empty(R, 3 * d, device=dev, dtype=BF)
        s["dq_acc"] = torch.randn(Bn, T, d, device=dev, generator=g)
        s["dw"] = torch.zeros(2, 64, device=dev)
        s["h12"] = torch.randn(R, 8 * d, device=dev, generator=g).bfloat16()
        s["dact"] = torch.randn(R, 4 * d, device=dev, generator=g).bfloat16()
        s["db"] = torch.zeros(8 * d, device=dev)
    add("qknorm_rope_fwd", 4 * per, lambda i: ops.qknorm_rope_fwd(S(i)["qkv"], wq, wk, rope, d, T))
    add("qknorm_rope_bwd (bf16 dq)", 6 * per, lambda i: ops.qknorm_rope_bwd(
        S(i)["dqk"], S(i)["qkv"], wq, wk, rope, S(i)["dqkv"], S(i)["dw"][0], S(i)["dw"][1], d, T))
    add("qknorm_rope_bwd (fp32 dq_acc)", 7 * per, lambda i: ops.qknorm_rope_bwd(
        S(i)["dqk"], S(i)["qkv"], wq, wk, rope, S(i)["dqkv"], S(i)["dw"][0], S(i)["dw"][1], d, T, dq_acc=S(i)["dq_acc"]))
    add("swiglu_bwd", 20 * per, lambda i: ops.swiglu_bwd(S(i)["dact"], S(i)["h12"], S(i)["db"]))
    print(f"[perf {name}] B={Bn} rows/sample={T} d={d}  ({nbuf} rotating buffer sets, {per / 1e6:.1f} MB per tensor)")
    print(f"    {'':30s} {'generation 1':>34s}   {'generation 2':>34s}")
    for tag, us, gbs in rows:
        cells = [f"{u:8.1f} us {g:7.0f} GB/s {g / HBM_PEAK:5.2f} of peak" for u, g in zip(us, gbs)]
        print(f"    {tag:30s} {cells[0]}   {cells[1]}")


def check_all():
    """fused gate+LN forward == the two kernels, and every second-generation kernel (the cluster fold of the
    column sums included: cluster sizes 1, 5 and 6 occur below) == its first-generation counterpart"""
    global ok_all
    ok_all = True
    try:
        for shp in [(2, 256, 256), (3, 77, 1216), (2, 64, 1536), (5, 154, 768), (3, 1, 128), (64, 256, 768), (16, 1024, 1536)]:
            check(*shp)
        for shp in [(2, 256, 256), (3, 77, 1216), (2, 64, 1536), (5, 154, 768), (3, 1, 128), (7, 33, 64),
                    (64, 256, 768), (64, 154, 768), (16, 1024, 1536), (200, 3, 512)]:
            check_generations(*shp)
    finally:
        ops.set_row_kernel_generation(2)
    print("ROW CHECK", "PASS" if ok_all else "FAIL")
    return ok_all


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("check", "all"):
        check_all()
    if what in ("perf", "all"):
        perf("cfg2 image", 64, 256, 768)
        perf("cfg2 text", 64, 154, 768)
        perf("cfg3 image", 64, 256, 1536)
        perf("cfg4 image", 16, 1024, 1536)
    sys.exit(0 if ok_all else 1)

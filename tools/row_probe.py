"""GB/s table of the LayerNorm-modulate / gate row kernels at the shapes the train step runs them at
(CUDA-graph replay over rotating buffer sets larger than the L2), plus a bit-for-bit check of the fused
gate+residual+LN forward against the two kernels it replaces.
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
        us = timeit(fn, nbuf)
        rows.append((tag, us, nbytes / us / 1e3))

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
    for tag, us, gbs in rows:
        print(f"    {tag:30s} {us:8.1f} us  {gbs:7.0f} GB/s  {gbs / HBM_PEAK:5.2f} of HBM peak")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("check", "all"):
        for shp in [(2, 256, 256), (3, 77, 1216), (2, 64, 1536), (5, 154, 768), (3, 1, 128), (64, 256, 768), (16, 1024, 1536)]:
            check(*shp)
        print("ROW CHECK", "PASS" if ok_all else "FAIL")
    if what in ("perf", "all"):
        perf("cfg2 image", 64, 256, 768)
        perf("cfg2 text", 64, 154, 768)
        perf("cfg3 image", 64, 256, 1536)
        perf("cfg4 image", 16, 1024, 1536)
    sys.exit(0 if ok_all else 1)

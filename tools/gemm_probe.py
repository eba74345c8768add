"""GPU diagnostic for the tcgen05 GEMM: compares against torch fp32 matmul on
the same bf16 inputs, case by case, and prints an error map when a case is off.
Run on the B200 box:  python tools/gemm_probe.py [group]
"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib, ops  # noqa: E402

L = _lib.lib()
dev = "cuda"


def run_gemm(A, B, M, N, K, a_major=0, b_major=0, d_fp32=False, accumulate=False, split_k=0,
             epilogue=0, bias=None, gate=None, rows_per_gate=0, resid=None, aux=False,
             remap=None, block_n=0, D=None, simt=False, out_rows=None):
    rows = out_rows if out_rows is not None else M
    if D is None:
        D = torch.zeros(rows, N, device=dev, dtype=torch.float32 if d_fp32 else torch.bfloat16)
    auxT = torch.zeros(rows, N, device=dev, dtype=torch.bfloat16) if aux else None
    a = _lib.GemmArgs()
    a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldb, a.ldd = A.stride(0), B.stride(0), D.stride(0)
    a.a_major, a.b_major = a_major, b_major
    a.d_fp32, a.accumulate, a.split_k, a.epilogue = int(d_fp32), int(accumulate), split_k, epilogue
    if bias is not None:
        a.bias, a.bias_fp32 = bias.data_ptr(), int(bias.dtype == torch.float32)
    if gate is not None:
        a.gate, a.rows_per_gate, a.ld_gate = gate.data_ptr(), rows_per_gate, gate.stride(0)
    if resid is not None:
        a.resid, a.ldr = resid.data_ptr(), resid.stride(0)
    if auxT is not None:
        a.aux, a.ld_aux = auxT.data_ptr(), auxT.stride(0)
    if remap is not None:
        a.remap_rows, a.remap_batch_rows, a.remap_offset = remap
    a.force_block_n = block_n
    fn = L.mmdit_gemm_bf16_simt if simt else L.mmdit_gemm_bf16
    rc = fn(C.byref(a), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gemm")
    return D, auxT


def err_map(got, ref, tag):
    diff = (got.float() - ref.float()).abs()
    scale = ref.float().abs().max().item() + 1e-20
    rel = diff.max().item() / scale
    print(f"    {tag}: max_abs={diff.max().item():.4e} ref_max={scale:.4e} rel={rel:.3e}")
    if rel > 2e-2:
        M, N = diff.shape
        bad = diff > 2e-2 * scale
        print(f"    bad fraction {bad.float().mean().item():.4f}; bad rows (first 16 of 8-row groups):",
              [i for i in range(0, M, 8) if bad[i:i + 8].any()][:16])
        print("    bad col groups (of 8):", [j for j in range(0, N, 8) if bad[:, j:j + 8].any()][:32])
        r0 = int(torch.nonzero(bad)[0][0]); c0 = int(torch.nonzero(bad)[0][1])
        print(f"    first bad at ({r0},{c0}) got={got[r0, c0].item():.5f} ref={ref[r0, c0].item():.5f}")
        print("    got[0,:8]", got[0, :8].float().tolist())
        print("    ref[0,:8]", ref[0, :8].float().tolist())
    return rel


def case(name, M, N, K, a_major=0, b_major=0, **kw):
    torch.manual_seed(hash(name) % 1000)
    A = (torch.randn(K, M, device=dev) if a_major else torch.randn(M, K, device=dev)).bfloat16()
    B = (torch.randn(K, N, device=dev) if b_major else torch.randn(N, K, device=dev)).bfloat16()
    Af = A.float().t() if a_major else A.float()
    Bf = B.float().t() if b_major else B.float()
    ref = Af @ Bf.t()
    print(f"[{name}] M={M} N={N} K={K} a_major={a_major} b_major={b_major} {kw}")
    try:
        D, _ = run_gemm(A, B, M, N, K, a_major, b_major, **kw)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("    EXCEPTION:", e)
        return False
    rel = err_map(D, ref, "tcgen05 vs torch")
    ok = rel < 1e-2
    print("    ->", "PASS" if ok else "FAIL")
    return ok


def group_basic():
    ok = True
    ok &= case("k64_n256", 128, 256, 64)
    ok &= case("k64_n128", 128, 128, 64, block_n=128)
    ok &= case("k64_n64", 128, 64, 64, block_n=64)
    ok &= case("k128", 256, 256, 128)
    ok &= case("k768", 512, 768, 768)
    ok &= case("ragged", 308, 768, 256)
    ok &= case("ragged_n", 308, 200, 80)
    ok &= case("k16", 512, 256, 16)
    ok &= case("big", 4096, 2304, 768)
    return ok


def group_major():
    ok = True
    ok &= case("dgrad_small", 128, 256, 64, 0, 1)
    ok &= case("dgrad", 512, 768, 3072, 0, 1)
    ok &= case("dgrad_n128", 512, 128, 256, 0, 1, block_n=128)
    ok &= case("amn_small", 128, 256, 64, 1, 0)
    ok &= case("wgrad_small", 128, 256, 64, 1, 1, d_fp32=True)
    ok &= case("wgrad", 768, 768, 4096, 1, 1, d_fp32=True)
    ok &= case("wgrad_ragged", 768, 3072, 308, 1, 1, d_fp32=True)
    ok &= case("wgrad_splitk", 768, 768, 16384, 1, 1, d_fp32=True, accumulate=True)
    ok &= case("wgrad_splitk4", 256, 256, 4096, 1, 1, d_fp32=True, accumulate=True, split_k=4)
    ok &= case("wgrad_n16", 64, 768, 2048, 1, 1, d_fp32=True)
    return ok


def group_epi():
    ok = True
    M, N, K = 616, 768, 256
    torch.manual_seed(3)
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    rows_per_gate = 154
    gate = torch.randn(M // rows_per_gate, N, device=dev).bfloat16()
    resid = torch.randn(M, N, device=dev).bfloat16()
    base = A.float() @ B.float().t() + bias
    for epi, name in [(0, "bias"), (1, "gate_resid"), (2, "silu"), (3, "resid")]:
        if epi == 0:
            ref = base
        elif epi == 1:
            ref = base * gate.float().repeat_interleave(rows_per_gate, 0) + resid.float()
        elif epi == 2:
            ref = torch.nn.functional.silu(base)
        else:
            ref = base + resid.float()
        print(f"[epi_{name}]")
        D, auxT = run_gemm(A, B, M, N, K, epilogue=epi, bias=bias, gate=gate,
                           rows_per_gate=rows_per_gate, resid=resid, aux=True)
        Ds, _ = run_gemm(A, B, M, N, K, epilogue=epi, bias=bias, gate=gate,
                         rows_per_gate=rows_per_gate, resid=resid, simt=True)
        torch.cuda.synchronize()
        r1 = err_map(D, ref, "tcgen05 vs torch")
        r2 = err_map(Ds, ref, "simt vs torch")
        r3 = err_map(auxT, base, "aux vs torch")
        ok &= r1 < 1e-2 and r2 < 1e-2 and r3 < 1e-2
    # bf16 bias + remap: scatter 77-row groups into 154-row batches at offset 77
    print("[remap]")
    Bt = 4
    A2 = torch.randn(Bt * 77, K, device=dev).bfloat16()
    ref2 = A2.float() @ B.float().t()
    D = torch.zeros(Bt * 154, N, device=dev, dtype=torch.bfloat16)
    run_gemm(A2, B, Bt * 77, N, K, remap=(77, 154, 77), D=D)
    torch.cuda.synchronize()
    got = D.view(Bt, 154, N)[:, 77:].reshape(Bt * 77, N)
    ok &= err_map(got, ref2, "remap rows") < 1e-2
    ok &= float(D.view(Bt, 154, N)[:, :77].abs().max()) == 0.0
    print("    ->", "PASS" if ok else "FAIL")
    return ok


def group_swiglu():
    """Fused SwiGLU epilogue == plain GEMM followed by mmdit_swiglu_fwd, bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
    from mmdit import ops
    ok = True
    for (M, d) in [(616, 256), (16384, 768), (9856, 768), (300, 128)]:
        torch.manual_seed(M)
        x = torch.randn(M, d, device=dev).bfloat16()
        w = (torch.randn(8 * d, d, device=dev) / d ** 0.5).bfloat16()
        b = torch.randn(8 * d, device=dev)
        h_ref = ops.gemm(x, w, bias=b)
        a_ref = ops.swiglu_fwd(h_ref)
        h = torch.full((M, 8 * d), float("nan"), device=dev, dtype=torch.bfloat16)
        a = ops.gemm(x, w, bias=b, epilogue=ops.EPI_SWIGLU, aux=h)
        torch.cuda.synchronize()
        same_h = bool((h == h_ref).all()) and not bool(torch.isnan(h.float()).any())
        same_a = bool((a == a_ref).all())
        t_ref = h_ref.float()
        tor = torch.nn.functional.silu(t_ref[:, :4 * d]) * t_ref[:, 4 * d:]
        rel = float((a.float() - tor).abs().max() / tor.abs().max())
        print(f"[swiglu M={M} d={d}] pre-activation identical: {same_h}, activation identical: {same_a}, vs torch rel {rel:.2e}")
        ok &= same_h and same_a and rel < 1e-2
    for (M, d) in [(16384, 768), (9856, 768), (16384, 1536)]:
        x = torch.randn(M, d, device=dev).bfloat16(); w = torch.randn(8 * d, d, device=dev).bfloat16(); b = torch.randn(8 * d, device=dev)
        h = torch.empty((M, 8 * d), device=dev, dtype=torch.bfloat16)
        def t(fn, iters=10):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters): fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        fused = t(lambda: ops.gemm(x, w, bias=b, epilogue=ops.EPI_SWIGLU, aux=h))
        plain = t(lambda: ops.gemm(x, w, bias=b))
        both = t(lambda: ops.swiglu_fwd(ops.gemm(x, w, bias=b)))
        print(f"[swiglu perf M={M} d={d}] fused {fused:.1f} us | plain GEMM {plain:.1f} us | GEMM + activation kernel {both:.1f} us")
    return ok


def group_qknorm():
    """EXPERIMENTAL epilogue (MMDIT_EPI_QKNORM): fused q|k|v projection + per-head RMSNorm + 2-D RoPE
    == plain GEMM followed by mmdit_qknorm_rope_fwd, bit for bit (image stream with RoPE tables and
    text stream without)."""
    sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
    from mmdit import ops
    ok = True
    for (B, T, d, rope) in [(2, 256, 256, True), (2, 154, 256, False), (64, 256, 768, True), (64, 154, 768, False),
                            (3, 240, 128, True)]:
        torch.manual_seed(T + d)
        R = B * T
        x = torch.randn(R, d, device=dev).bfloat16()
        w = (torch.randn(3 * d, d, device=dev) / d ** 0.5).bfloat16()
        wq = (1 + 0.1 * torch.randn(64, device=dev)).float(); wk = (1 + 0.1 * torch.randn(64, device=dev)).float()
        tabs = None
        if rope:
            ang = torch.rand(T, 32, device=dev) * 6.28
            tabs = (torch.cos(ang).contiguous(), torch.sin(ang).contiguous())
        qkv_ref = ops.gemm(x, w)
        qk_ref = ops.qknorm_rope_fwd(qkv_ref, wq, wk, tabs, d, T)
        qk = torch.full((R, 2 * d), float("nan"), device=dev, dtype=torch.bfloat16)
        cos, sin = tabs if tabs is not None else (None, None)
        qkv = ops.gemm(x, w, epilogue=ops.EPI_QKNORM, aux=qk, qknorm=(wq, wk, cos, sin, T))
        torch.cuda.synchronize()
        same_raw = bool((qkv == qkv_ref).all())
        same_qk = bool((qk == qk_ref).all())
        rel = float((qk.float() - qk_ref.float()).abs().max() / qk_ref.float().abs().max())
        print(f"[qknorm B={B} T={T} d={d} rope={rope}] raw identical: {same_raw}, q|k identical: {same_qk} (max rel {rel:.2e})")
        ok &= same_raw and rel < 1e-2
    return ok


def group_swiglu_bwd():
    """w3 data-gradient GEMM with the SwiGLU backward in its epilogue vs the two-kernel path (bit-identical
    dh12, bias-gradient column sums), and both timed under a CUDA graph."""
    ok = True
    for (R, d, hid) in [(512, 256, 1024), (1280, 384, 768), (16384, 768, 3072), (9856, 768, 3072), (16384, 1536, 6144)]:
        g = torch.Generator(device=dev).manual_seed(R + d)
        dy = (0.5 * torch.randn(R, d, device=dev, generator=g)).bfloat16()
        w3 = (torch.randn(d, hid, device=dev, generator=g) / d ** 0.5).bfloat16()
        h12 = torch.randn(R, 2 * hid, device=dev, generator=g).bfloat16()
        db_f = torch.zeros(2 * hid, device=dev)
        dh_f = ops.gemm_swiglu_bwd(dy, w3, h12, db_f)
        da = ops.gemm(dy, w3, b_major=1)
        db_u = torch.zeros(2 * hid, device=dev)
        dh_u = ops.swiglu_bwd(da, h12, db_u)
        torch.cuda.synchronize()
        ndiff = int((dh_f.view(torch.int16) != dh_u.view(torch.int16)).sum())
        rel_db = float((db_f - db_u).abs().max() / db_u.abs().max())
        nob = ops.gemm_swiglu_bwd(dy, w3, h12, None)
        same_nob = bool(torch.equal(nob, dh_f))
        if ndiff:
            bad = torch.nonzero(dh_f.view(torch.int16) != dh_u.view(torch.int16))
            rows = bad[:, 0].unique()
            print("    bad rows", rows[:12].tolist(), "n_rows", rows.numel(), "cols of first row",
                  bad[bad[:, 0] == rows[0], 1].tolist()[:40])
            r, c = int(bad[0, 0]), int(bad[0, 1])
            print("    first bad: fused", float(dh_f[r, c]), "unfused", float(dh_u[r, c]), "row%128", r % 128, "col%64", c % 64, "half", c // hid)
            again = ops.gemm_swiglu_bwd(dy, w3, h12, torch.zeros_like(db_f))
            print("    fused run twice identical:", bool(torch.equal(again, dh_f)),
                  " second run vs unfused:", int((again.view(torch.int16) != dh_u.view(torch.int16)).sum()))
        if R >= 8192 and os.environ.get("SWIGLU_BWD_DEBUG"):
            for dbg in (0,):
                bad_runs = []
                for rep in range(4):
                    out = ops.gemm_swiglu_bwd(dy, w3, h12, None, debug=dbg)
                    bad_runs.append(int((out.view(torch.int16) != dh_u.view(torch.int16)).sum()))
                print(f"    debug={dbg:3d}: mismatches per run {bad_runs}")
        good = ndiff == 0 and rel_db < 2e-3 and same_nob
        ok &= good
        print(f"[swiglu_bwd R={R} d={d} hid={hid}] differing dh12 elements: {ndiff} of {dh_f.numel()}, "
              f"db max-rel {rel_db:.2e}, without column sums identical: {same_nob}  {'ok' if good else 'FAIL'}")
        if R >= 8192:
            def timeit(fn, n=8):
                fn(); torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                st = torch.cuda.Stream()
                with torch.cuda.stream(st):
                    with torch.cuda.graph(gr, stream=st):
                        for _ in range(n):
                            fn()
                gr.replay(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    gr.replay()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / (5 * n) * 1e3
            dbz = torch.zeros(2 * hid, device=dev)
            t_f = timeit(lambda: ops.gemm_swiglu_bwd(dy, w3, h12, dbz))
            t_g = timeit(lambda: ops.gemm(dy, w3, b_major=1))
            t_s = timeit(lambda: ops.swiglu_bwd(da, h12, dbz))
            print(f"    fused {t_f:7.1f} us   vs   GEMM {t_g:6.1f} + swiglu_bwd {t_s:6.1f} = {t_g + t_s:6.1f} us")
    return ok


def group_perf():
    shapes = [
        ("qkv_x cfg2", 16384, 2304, 768, 0, 0, False),
        ("w12_x cfg2", 16384, 6144, 768, 0, 0, False),
        ("w3_x cfg2", 16384, 768, 3072, 0, 0, False),
        ("w12 dgrad cfg2", 16384, 768, 6144, 0, 1, False),
        ("w12 wgrad cfg2", 6144, 768, 16384, 1, 1, True),
        ("out wgrad cfg2", 768, 768, 16384, 1, 1, True),
        ("w12_x cfg3", 16384, 12288, 1536, 0, 0, False),
        ("sq 8192", 8192, 8192, 8192, 0, 0, False),
    ]
    for name, M, N, K, am, bm, f32 in shapes:
        A = (torch.randn(K, M, device=dev) if am else torch.randn(M, K, device=dev)).bfloat16()
        B = (torch.randn(K, N, device=dev) if bm else torch.randn(N, K, device=dev)).bfloat16()
        D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        for bn in (0, 128, 256):
            for _ in range(3):
                run_gemm(A, B, M, N, K, am, bm, d_fp32=f32, accumulate=f32, D=D, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                run_gemm(A, B, M, N, K, am, bm, d_fp32=f32, accumulate=f32, D=D, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"[perf] {name:18s} bn={bn:3d} {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
        # torch reference speed
        Af = A.t() if am else A
        Bf = B.t() if bm else B
        for _ in range(3):
            torch.matmul(Af, Bf.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(Af, Bf.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"[perf] {name:18s} cuBLAS {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
    return True


if __name__ == "__main__":
    g = sys.argv[1] if len(sys.argv) > 1 else "basic"
    _lib.check(L.mmdit_device_check(), "device_check")
    t0 = time.time()
    ok = {"basic": group_basic, "major": group_major, "epi": group_epi, "perf": group_perf, "swiglu": group_swiglu, "qknorm": group_qknorm, "swiglu_bwd": group_swiglu_bwd}[g]()
    print(f"GROUP {g}: {'ALL PASS' if ok else 'SOME FAIL'} ({time.time() - t0:.1f}s)")
    sys.exit(0 if ok else 1)

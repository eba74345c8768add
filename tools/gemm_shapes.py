"""Every distinct GEMM of the cfg2 (or cfg3) train step, timed alone: our tcgen05 kernel (through
ops.gemm, i.e. including the split-K fold where the wrapper plans one) next to cuBLAS (torch.matmul)
on the same operands.  Inputs rotate over several buffers so that successive calls do not hit in L2.
usage: python tools/gemm_shapes.py [cfg2|cfg3]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
d = 768 if which == "cfg2" else 1536
depth = 12 if which == "cfg2" else 24
RX, RC = 64 * 256, 64 * 154
shapes = []  # (name, M, N, K, a_major, b_major, f32out, bias, count per step)
for tag, R, cnt in (("x", RX, depth), ("c", RC, depth - 1)):
    shapes += [
        (f"qkv_{tag} fprop", R, 3 * d, d, 0, 0, 0, 0, cnt),
        (f"out_{tag} fprop", R, d, d, 0, 0, 0, 0, cnt),
        (f"w12_{tag} fprop", R, 8 * d, d, 0, 0, 0, 1, cnt),
        (f"w3_{tag} fprop", R, d, 4 * d, 0, 0, 0, 1, cnt),
        (f"qkv_{tag} dgrad", R, d, 3 * d, 0, 1, 0, 0, cnt),
        (f"out_{tag} dgrad", R, d, d, 0, 1, 0, 0, cnt),
        (f"w12_{tag} dgrad", R, d, 8 * d, 0, 1, 0, 0, cnt),
        (f"w3_{tag} dgrad", R, 4 * d, d, 0, 1, 0, 0, cnt),
        (f"qkv_{tag} wgrad", 3 * d, d, R, 1, 1, 1, 0, cnt),
        (f"out_{tag} wgrad", d, d, R, 1, 1, 1, 0, cnt),
        (f"w12_{tag} wgrad", 8 * d, d, R, 1, 1, 1, 0, cnt),
        (f"w3_{tag} wgrad", d, 4 * d, R, 1, 1, 1, 0, cnt),
    ]
shapes += [
    ("adaLN fprop", 64, 12 * d, d, 0, 0, 0, 0, depth),
    ("adaLN dgrad", 64, d, 12 * d, 0, 1, 0, 0, depth),
    ("adaLN wgrad", 12 * d, d, 64, 1, 1, 1, 0, depth),
]


def bench(fn, iters=12):
    """Device time per call: the calls are captured into one CUDA graph (no host launch gaps)."""
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


NB = 3
tot_o = tot_c = 0.0
print(f"{'gemm':16s} {'M':>6} {'N':>6} {'K':>6} aM bM out | ours us   TF/s | cuBLAS us  TF/s | ratio | per-step ms ours/cublas")
for name, M, N, K, am, bm, f32, bias, cnt in shapes:
    As = [torch.randn((K, M) if am else (M, K), device=dev).bfloat16() for _ in range(NB)]
    Bs = [torch.randn((K, N) if bm else (N, K), device=dev).bfloat16() for _ in range(NB)]
    bv = torch.randn(N, device=dev) if bias else None
    odt = torch.float32 if f32 else torch.bfloat16

    def ours(i):
        return ops.gemm(As[i % NB], Bs[i % NB], a_major=am, b_major=bm, out_dtype=odt, bias=bv)

    def cub(i):
        A = As[i % NB].t() if am else As[i % NB]
        B = Bs[i % NB] if bm else Bs[i % NB].t()
        y = torch.matmul(A, B)
        return y

    ref = cub(0).float() + (bv if bias else 0)
    got = ours(0).float()
    err = float((got - ref).abs().max() / ref.abs().max())
    uo, uc = bench(ours), bench(cub)
    fl = 2.0 * M * N * K
    tot_o += uo * cnt
    tot_c += uc * cnt
    print(f"{name:16s} {M:6d} {N:6d} {K:6d} {am:2d} {bm:2d} {'f32' if f32 else 'b16'} | {uo:7.1f} {fl/uo/1e6:7.0f} | "
          f"{uc:7.1f} {fl/uc/1e6:7.0f} | {uo/uc:5.2f} | {uo*cnt/1e3:6.2f} {uc*cnt/1e3:6.2f}  err {err:.1e}")
print(f"per-step GEMM time: ours {tot_o/1e3:.2f} ms, cuBLAS (bf16 out, no fp32 wgrad / bias) {tot_c/1e3:.2f} ms")

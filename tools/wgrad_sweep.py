"""Small-output wgrad GEMM sweep (MN-major x MN-major, long K): split-K / tile-shape experiments."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib
L = _lib.lib(); dev = "cuda"
def run(M, N, K, split, bn, iters=10):
    A = torch.randn(K, M, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
    D = torch.zeros(max(split, 1) * M, N, device=dev, dtype=torch.float32)
    a = _lib.GemmArgs(); a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    a.M, a.N, a.K = M, N, K; a.lda, a.ldb, a.ldd = M, N, N; a.a_major = a.b_major = 1
    a.d_fp32 = 1; a.split_k = split; a.force_block_n = bn
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): _lib.check(L.mmdit_gemm_bf16(C.byref(a), s), "gemm")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.mmdit_gemm_bf16(C.byref(a), s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (M, N, K) in [(768, 3072, 16384), (3072, 768, 16384), (2304, 768, 16384), (768, 768, 16384), (6144, 768, 16384)]:
    out = []
    for bn in (256, 128):
        for split in (1, 2, 3, 4, 8):
            us = run(M, N, K, split, bn)
            out.append(f"bn{bn}/s{split}: {us:6.1f}us {2.0*M*N*K/us/1e6:6.0f}TF")
    print(f"M={M} N={N} K={K}:\n   " + "\n   ".join(" | ".join(out[i:i+5]) for i in (0, 5)))

import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib
L = _lib.lib(); dev = "cuda"
def run(M, N, K, epi, debug=0, bn=0, iters=20):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16); aux = torch.empty_like(D)
    resid = torch.randn(M, N, device=dev).bfloat16(); gate = torch.randn(M // 256, N, device=dev).bfloat16()
    a = _lib.GemmArgs(); a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    a.M, a.N, a.K = M, N, K; a.lda, a.ldb, a.ldd = K, K, N; a.reserved = debug; a.force_block_n = bn
    a.epilogue = epi
    if epi == 1:
        a.gate, a.rows_per_gate, a.ld_gate = gate.data_ptr(), 256, N
        a.resid, a.ldr = resid.data_ptr(), N
        a.aux, a.ld_aux = aux.data_ptr(), N
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): _lib.check(L.mmdit_gemm_bf16(C.byref(a), s), "gemm")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.mmdit_gemm_bf16(C.byref(a), s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (M, N, K) in [(16384, 768, 768), (16384, 768, 3072)]:
    for bn in (256, 128):
        print(f"M={M} N={N} K={K} bn={bn}: plain {run(M,N,K,0,0,bn):.1f}us | gate {run(M,N,K,1,0,bn):.1f} | no-store {run(M,N,K,1,1,bn):.1f}"
              f" | no-aux {run(M,N,K,1,4,bn):.1f} | no-prefetch {run(M,N,K,1,8,bn):.1f} | no-aux,no-pref {run(M,N,K,1,12,bn):.1f}")

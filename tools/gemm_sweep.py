"""K sweep + epilogue ablations for the tcgen05 GEMM (perf experiments)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib
L = _lib.lib(); dev = "cuda"

def run(M, N, K, debug=0, bn=0, iters=20):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    a = _lib.GemmArgs(); a.A, a.B, a.D = A.data_ptr(), B.data_ptr(), D.data_ptr()
    a.M, a.N, a.K = M, N, K; a.lda, a.ldb, a.ldd = K, K, N; a.reserved = debug; a.force_block_n = bn
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): _lib.check(L.mmdit_gemm_bf16(C.byref(a), s), "gemm")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.mmdit_gemm_bf16(C.byref(a), s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

if __name__ == '__main__' and len(sys.argv) > 1:
    M, N, K = (int(x) for x in sys.argv[1:4])
    dbg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    print(M, N, K, dbg, bn, run(M, N, K, dbg, bn, iters=3), 'us')
    sys.exit(0)
M, N = 16384, 6144
waves = -(-(M // 128) * (N // 256) // 148)
for K in (64, 256, 768, 1536, 3072, 6144):
    row = []
    for dbg in (0, 1, 2, 3):
        us = run(M, N, K, dbg)
        row.append(f"dbg{dbg} {us:8.1f}us {2.0*M*N*K/us/1e6:7.1f}TF")
    print(f"K={K:5d} kb={K//64:3d} | " + " | ".join(row) + f" | per-tile {run(M,N,K)/waves:.2f}us")

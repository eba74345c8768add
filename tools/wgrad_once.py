"""Runs each small-output wgrad GEMM of cfg2 a few times (eager) -- meant to be run under
ncu --metrics gpu__time_duration.sum to see GEMM and fold kernels separately."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops
dev = "cuda"
for (M, N, K) in [(2304, 768, 16384), (768, 3072, 16384), (768, 768, 16384), (6144, 768, 16384),
                  (2304, 768, 9856), (768, 3072, 9856)]:
    A = torch.randn(K, M, device=dev).bfloat16(); B = torch.randn(K, N, device=dev).bfloat16()
    for _ in range(3):
        ops.gemm(A, B, a_major=1, b_major=1, out_dtype=torch.float32)
    torch.cuda.synchronize()
    print(M, N, K, "split", ops._plan_split(M, N, K))

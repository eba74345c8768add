"""Repro helper: JointAttentionFn.forward-like call at a given shape (strided q/k from the qk buffers)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops
B, H, N, M, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
d = H * 64
dev = "cuda"
torch.manual_seed(0)
qkv_x = torch.randn(B * N, 3 * d, device=dev).bfloat16(); qkv_c = torch.randn(B * M, 3 * d, device=dev).bfloat16()
w = torch.ones(64, device=dev)
if mode.startswith("qk"):
    qk_x = ops.qknorm_rope_fwd(qkv_x, w, w, None, d, N); qk_c = ops.qknorm_rope_fwd(qkv_c, w, w, None, d, M)
    q = (qk_x[:, :d], qk_c[:, :d]); k = (qk_x[:, d:], qk_c[:, d:])
else:
    q = (qkv_x[:, :d], qkv_c[:, :d]); k = (qkv_x[:, d:2 * d], qkv_c[:, d:2 * d])
v = (qkv_x[:, 2 * d:], qkv_c[:, 2 * d:])
bound = ops.qk_logit_bound(w, w, w, w, 0.125) if mode.endswith("bound") else None
torch.cuda.synchronize()
o_x, o_c, lse = ops.attn_fwd(q, k, v, B, H, N, M, 0.125, logit_bound=bound)
torch.cuda.synchronize()
print(f"OK B={B} H={H} N={N} M={M} {mode}: finite {bool(torch.isfinite(o_x.float()).all())} {bool(torch.isfinite(lse).all())}")

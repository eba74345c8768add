"""Launch the attention kernels a few times at one bench shape (for ncu captures).
usage: python tools/attn_one.py [cfg2|cfg4] [iters]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, H, N, M = {"cfg2": (64, 12, 256, 154), "cfg3": (64, 24, 256, 154), "cfg4": (16, 24, 1024, 154)}[which]
dev, d = "cuda", H * 64
torch.manual_seed(0)


def unit(t):
    qk = t[:, :2 * d].float().view(t.shape[0], 2 * H, 64)
    t[:, :2 * d] = (qk * torch.rsqrt(qk.pow(2).mean(-1, keepdim=True))).view(t.shape[0], 2 * d).bfloat16()
    return t


qkv_x = unit(torch.randn(B * N, 3 * d, device=dev).bfloat16())
qkv_c = unit(torch.randn(B * M, 3 * d, device=dev).bfloat16())
one = torch.ones(64, device=dev)
bound = ops.qk_logit_bound(one, one, one, one, 0.125)
qs, ks, vs = ((qkv_x[:, i * d:(i + 1) * d], qkv_c[:, i * d:(i + 1) * d]) for i in range(3))
dx, dc = torch.empty_like(qkv_x), torch.empty_like(qkv_c)
dq, dk, dv = ((dx[:, i * d:(i + 1) * d], dc[:, i * d:(i + 1) * d]) for i in range(3))
for _ in range(iters):
    o_x, o_c, lse = ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound)
    do_x, do_c = torch.randn_like(o_x), torch.randn_like(o_c)
    ops.attn_bwd(qs, ks, vs, (o_x, o_c), lse, (do_x, do_c), dq, dk, dv, B, H, N, M, 0.125)
torch.cuda.synchronize()
print("done", float(o_x.float().abs().mean()), float(dx.float().abs().mean()))

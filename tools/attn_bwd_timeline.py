"""Dump the in-kernel timeline of one attention-backward CTA (tuning)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib, ops
L = _lib.lib()
B, H, N, M = 64, 12, 256, 154
d = H * 64
dev = "cuda"
qkv_x = torch.randn(B * N, 3 * d, device=dev).bfloat16(); qkv_c = torch.randn(B * M, 3 * d, device=dev).bfloat16()
qs, ks, vs = ((qkv_x[:, i * d:(i + 1) * d], qkv_c[:, i * d:(i + 1) * d]) for i in range(3))
one = torch.ones(64, device=dev)
bound = ops.qk_logit_bound(one, one, one, one, 0.125) * 2.9
o_x, o_c, lse = ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound)
do_x, do_c = torch.randn_like(o_x), torch.randn_like(o_c)
dx, dc = torch.empty_like(qkv_x), torch.empty_like(qkv_c)
dq, dk, dv = ((dx[:, i * d:(i + 1) * d], dc[:, i * d:(i + 1) * d]) for i in range(3))
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
L.mmdit_debug_attn_bwd_timeline.argtypes = [C.c_void_p, C.c_int]
for _ in range(2):
    ops.attn_bwd(qs, ks, vs, (o_x, o_c), lse, (do_x, do_c), dq, dk, dv, B, H, N, M, 0.125)
for blk in [1500, 1501, 1503]:
    buf.zero_()
    assert L.mmdit_debug_attn_bwd_timeline(buf.data_ptr(), blk) == 0
    ops.attn_bwd(qs, ks, vs, (o_x, o_c), lse, (do_x, do_c), dq, dk, dv, B, H, N, M, 0.125)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    ev = []
    for base, who in ((0, "mma"), (256, "cmp")):
        for i in range(128):
            if base + 2 * i + 1 < len(t) and t[base + 2 * i]:
                ev.append((t[base + 2 * i + 1], who, t[base + 2 * i]))
    ev.sort()
    t0 = ev[0][0]
    print(f"--- BWD block {blk} (kv tile {blk % 4}) : total {ev[-1][0] - t0} cycles")
    for who in ("mma", "cmp"):
        print(who, " ".join(f"{eid}@{tt - t0}" for tt, w, eid in ev if w == who))
L.mmdit_debug_attn_bwd_timeline(None, -1)

"""Dump the in-kernel timeline of one CTA of the second-generation attention forward (tuning).
usage: python tools/attn2_timeline.py [cfg2|cfg4] [block ...]"""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import _lib, ops
L = _lib.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B, H, N, M = {"cfg2": (64, 12, 256, 154), "cfg4": (16, 24, 1024, 154)}[which]
d = H * 64
dev = "cuda"
qkv_x = torch.randn(B * N, 3 * d, device=dev).bfloat16(); qkv_c = torch.randn(B * M, 3 * d, device=dev).bfloat16()
qs, ks, vs = ((qkv_x[:, i * d:(i + 1) * d], qkv_c[:, i * d:(i + 1) * d]) for i in range(3))
one = torch.ones(64, device=dev)
bound = ops.qk_logit_bound(one, one, one, one, 0.125) * 2.9   # random-normal q,k: looser bound
for _ in range(3):
    ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound)
torch.cuda.synchronize()
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
L.mmdit_debug_attn2_timeline.argtypes = [C.c_void_p, C.c_int]
for blk in [int(x) for x in (sys.argv[2:] or [0, 101])]:
    buf.zero_()
    assert L.mmdit_debug_attn2_timeline(buf.data_ptr(), blk) == 0
    ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    ev = []
    for base, cap, who in ((0, 128, "mma"), (256, 256, "sA"), (768, 256, "sB")):
        for i in range(cap):
            if t[base + 2 * i]:
                ev.append((t[base + 2 * i + 1], who, t[base + 2 * i]))
    ev.sort()
    t0 = ev[0][0]
    print(f"--- block {blk}: {len(ev)} events, span {ev[-1][0] - t0} cycles")
    for who in ("mma", "sA", "sB"):
        print(who, " ".join(f"{eid}@{tt - t0}" for tt, w, eid in ev if w == who))
L.mmdit_debug_attn2_timeline(None, -1)

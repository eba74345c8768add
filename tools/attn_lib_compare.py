"""Same-box attention table: our tcgen05 kernels vs flash_attn_func 2.8.3 (the reference's call,
Attention.py:293) vs torch SDPA (cuDNN / flash backends), forward and backward, at the cfg2 / cfg3 /
cfg4 joint-sequence shapes.  Device time from CUDA-graph replays (no host launch gaps); inputs rotate
over several buffers so that successive calls do not hit in L2.
usage: python tools/attn_lib_compare.py [out.json]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit import ops  # noqa: E402

dev = "cuda"
BF = torch.bfloat16
NB = 3


def bench(fn, iters=9):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def unit_rms(t, d, H):
    qk = t[:, :2 * d].float().view(t.shape[0], 2 * H, 64)
    qk = qk * torch.rsqrt(qk.pow(2).mean(-1, keepdim=True))
    t[:, :2 * d] = qk.view(t.shape[0], 2 * d).bfloat16()
    return t


GEN = None


def randn(*shape, dtype=torch.float32):
    """Own CUDA generator: flash-attn registers the default one with the graphs captured below,
    after which drawing from it outside a capture raises."""
    global GEN
    if GEN is None:
        GEN = torch.Generator(device=dev)
        GEN.manual_seed(0)
    return torch.randn(*shape, device=dev, generator=GEN).to(dtype)


def main():
    rows = []
    for name, B, H, N, M in [("cfg2", 64, 12, 256, 154), ("cfg3", 64, 24, 256, 154), ("cfg4", 16, 24, 1024, 154)]:
        d, T = H * 64, N + M
        flops = 4.0 * B * H * T * T * 64
        one = torch.ones(64, device=dev)
        bound = ops.qk_logit_bound(one, one, one, one, 0.125)
        sets = []
        for _ in range(NB):
            qkv_x = unit_rms(randn(B * N, 3 * d, dtype=BF), d, H)
            qkv_c = unit_rms(randn(B * M, 3 * d, dtype=BF), d, H)
            qs, ks, vs = ((qkv_x[:, i * d:(i + 1) * d], qkv_c[:, i * d:(i + 1) * d]) for i in range(3))
            o_x, o_c, lse = ops.attn_fwd(qs, ks, vs, B, H, N, M, 0.125, logit_bound=bound)
            do = (randn(*o_x.shape, dtype=BF), randn(*o_c.shape, dtype=BF))
            dx, dc = torch.empty_like(qkv_x), torch.empty_like(qkv_c)
            dq, dk, dv = ((dx[:, i * d:(i + 1) * d], dc[:, i * d:(i + 1) * d]) for i in range(3))
            sets.append((qs, ks, vs, (o_x, o_c), lse, do, dq, dk, dv))

        def ours_fwd(i):
            s = sets[i % NB]
            ops.attn_fwd(s[0], s[1], s[2], B, H, N, M, 0.125, logit_bound=bound)

        def ours_bwd(i):
            s = sets[i % NB]
            ops.attn_bwd(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], B, H, N, M, 0.125)

        r = {"shape": name, "B": B, "H": H, "T": T, "fwd_gflop": flops / 1e9}
        r["ours_fwd_us"] = bench(ours_fwd)
        r["ours_bwd_us"] = bench(ours_bwd)
        # library kernels on the joint sequence ([B, T, H, 64] for flash-attn, [B, H, T, 64] for SDPA)
        try:
            from flash_attn import flash_attn_func
            fa = [tuple(randn(B, T, H, 64, dtype=BF).requires_grad_(True) for _ in range(3)) for _ in range(NB)]
            r["fa2_fwd_us"] = bench(lambda i: flash_attn_func(*[t.detach() for t in fa[i % NB]], softmax_scale=0.125))
            outs = [flash_attn_func(*fa[i], softmax_scale=0.125) for i in range(NB)]
            gos = [randn(*o.shape, dtype=BF) for o in outs]
            r["fa2_bwd_us"] = bench(lambda i: torch.autograd.grad(outs[i % NB], fa[i % NB], gos[i % NB], retain_graph=True))
        except Exception as e:  # noqa: BLE001
            r["fa2_error"] = repr(e)[:200]
        sd = [tuple(randn(B, H, T, 64, dtype=BF).requires_grad_(True) for _ in range(3)) for _ in range(NB)]
        r["sdpa_fwd_us"] = bench(lambda i: F.scaled_dot_product_attention(*[t.detach() for t in sd[i % NB]], scale=0.125))
        try:
            outs2 = [F.scaled_dot_product_attention(*sd[i], scale=0.125) for i in range(NB)]
            gos2 = [randn(*o.shape, dtype=BF) for o in outs2]
            r["sdpa_bwd_us"] = bench(lambda i: torch.autograd.grad(outs2[i % NB], sd[i % NB], gos2[i % NB], retain_graph=True))
        except Exception as e:  # noqa: BLE001
            r["sdpa_bwd_error"] = repr(e)[:200]
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
                r["cudnn_fwd_us"] = bench(lambda i: F.scaled_dot_product_attention(*[t.detach() for t in sd[i % NB]], scale=0.125))
        except Exception as e:  # noqa: BLE001
            r["cudnn_error"] = repr(e)[:200]
        for k in list(r):
            if k.endswith("_fwd_us"):
                r[k.replace("_us", "_tflops")] = flops / r[k] / 1e6
            if k.endswith("_bwd_us"):
                r[k.replace("_us", "_tflops")] = 2.5 * flops / r[k] / 1e6
        rows.append(r)
        print(json.dumps(r))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

"""Debug: captured Euler step vs eager with checksum hooks (no big tensors retained)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from mmdit import ops
import mmdit.functional as Fn
from oracle import mmdit_oracle as O
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
m = diff_model(device=dev, **cfg).eval()
m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
m.load_text_encoders()
B, L, steps = 2, 32, 6
noise = torch.randn(B, 16, L, L).to(dev).float().contiguous()
th, tp = m.text_encoders.text_to_embedding("a prompt")
null = torch.tensor([0] * B + [1] * B).bool().to(dev)
th = th.repeat(2 * B, 1, 1).to(dev); tp = tp.repeat(2 * B, 1).to(dev)
dt = 1 / steps
sums = []          # (tag, scalar tensor)
mode = sys.argv[1] if len(sys.argv) > 1 else "ops"
if mode == "ops":
    # checksum every op output
    names = ["gemm", "attn_fwd", "ln_modulate_fwd", "gate_residual_fwd", "qknorm_rope_fwd", "text_norm_fwd",
             "patchify", "unpatchify", "timestep_embed_fwd", "qk_logit_bound"]
    for nme in names:
        orig = getattr(ops, nme)
        def wrap(*a, _o=orig, _n=nme, **k):
            out = _o(*a, **k)
            first = out[0] if isinstance(out, tuple) else out
            sums.append((_n, first.float().abs().sum()))
            return out
        setattr(ops, nme, wrap)
else:
    for i, blk in enumerate(m.blocks):
        blk.register_forward_hook(lambda mod, inp, out, i=i: sums.append((f"block{i}", out[0].float().abs().sum() + out[1].float().abs().sum())))
with torch.no_grad():
    xe = noise.clone()
    v = m.forward(xe.repeat(2, 1, 1, 1), torch.ones(2 * B, device=dev), th, tp, null, null, null)
    ops.cfg_euler_step(xe, v.contiguous(), 5.0, dt)
    eager = [(n, float(s)) for n, s in sums]; sums.clear()
    g, sx, st = m._euler_step_graph(noise.clone(), th, tp, null, 5.0, dt)
    per = len(eager)
    cap = sums[-per:]
    st.fill_(1.0); g.replay(); torch.cuda.synchronize()
    print("x after step 0: diff", float((sx - xe).abs().max()))
    bad = 0
    for (n, e), (n2, s) in zip(eager, cap):
        s = float(s)
        flag = "" if abs(s - e) <= 1e-6 * max(1.0, abs(e)) else "   <-- DIFFERS"
        if flag:
            bad += 1
        if flag or mode != "ops":
            print(f"   {n:20s} eager {e:.6e} graph {s:.6e}{flag}")
        if bad >= 6:
            break
    print("ops compared:", per, "first mismatches shown:", bad)

"""Debug: per-block outputs of the captured Euler step (built by diff_model._euler_step_graph) vs eager."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from mmdit import ops
from oracle import mmdit_oracle as O
dev = torch.device("cuda")
torch.manual_seed(0)
dim, heads, depth, L = 256, 4, 2, 32
pre_loop = len(sys.argv) > 1 and sys.argv[1] == "preloop"
cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=dim, hidden_scale=4.0, num_heads=heads,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=depth, positional_encoding="RoPE2d")
m = diff_model(device=dev, **cfg).eval()
m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
m.load_text_encoders()
B = 2
noise = torch.randn(B, 16, L, L).to(dev).float().contiguous()
th, tp = m.text_encoders.text_to_embedding("a prompt")
null = torch.tensor([0] * B + [1] * B).bool().to(dev)
th = th.repeat(2 * B, 1, 1).to(dev); tp = tp.repeat(2 * B, 1).to(dev)
t1 = torch.ones(2 * B, device=dev)
rec = []
hooks = [blk.register_forward_hook(lambda mod, inp, out: rec.append((out[0], out[1]))) for blk in m.blocks]
dt = 1 / 6
with torch.no_grad():
    if pre_loop:
        xe = noise.clone()
        for t in torch.linspace(1, dt, 6).to(dev):
            v = m.forward(xe.repeat(2, 1, 1, 1), t.repeat(2 * B), th, tp, null, null, null)
            ops.cfg_euler_step(xe, v.contiguous(), 5.0, dt)
        rec.clear()
    v_ref = m.forward(noise.repeat(2, 1, 1, 1), t1, th, tp, null, null, null).float().clone()
    eager = [(a.float().clone(), b.float().clone()) for a, b in rec]; rec.clear()
    x_ref = noise.clone(); ops.cfg_euler_step(x_ref, v_ref.bfloat16().contiguous(), 5.0, dt)
    g, sx, st = m._euler_step_graph(noise.clone(), th, tp, null, 5.0, dt)
    cap = rec[-depth:]          # the capture pass is the last forward recorded
    print("recorded forwards:", len(rec) // depth)
    st.copy_(t1); g.replay(); torch.cuda.synchronize()
    for i, ((a, b), (ea, eb)) in enumerate(zip(cap, eager)):
        print(f"   block {i}: x diff {float((a.float() - ea).abs().max()):.3e}  c diff {float((b.float() - eb).abs().max()):.3e}"
              f"   nan {int(torch.isnan(a.float()).sum())}")
    print("x after step 0: diff", float((sx - x_ref).abs().max()))
    # per-sample view of the first block's image-stream diff
    d = (cap[0][0].float() - eager[0][0]).abs().amax(dim=(1, 2))
    print("block 0 x diff per forward-batch row:", [f"{float(q):.2e}" for q in d])
    d = (cap[0][1].float() - eager[0][1]).abs().amax(dim=(1, 2))
    print("block 0 c diff per forward-batch row:", [f"{float(q):.2e}" for q in d])

"""Debug: captured Euler step vs eager Euler loop, step by step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from mmdit import ops
dev = torch.device("cuda")
torch.manual_seed(0)
dim, heads, depth, L, steps = (int(a) for a in (sys.argv[1:6] or [256, 4, 2, 16, 4]))
cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=dim, hidden_scale=4.0, num_heads=heads,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=depth, positional_encoding="RoPE2d")
m = diff_model(device=dev, **cfg).eval()
from oracle import mmdit_oracle as O
m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
m.load_text_encoders()
B = 2
noise = torch.randn(B, 16, L, L).to(dev).float().contiguous()
th, tp = m.text_encoders.text_to_embedding("a prompt")
null = torch.tensor([0] * B + [1] * B).bool().to(dev)
th = th.repeat(2 * B, 1, 1).to(dev); tp = tp.repeat(2 * B, 1).to(dev)
ts = torch.linspace(1, 1.0 / steps, steps).to(dev)
dt = 1 / steps
with torch.no_grad():
    xe = noise.clone(); eager_states = []
    for t in ts:
        v = m.forward(xe.repeat(2, 1, 1, 1), t.repeat(2 * B), th, tp, null, null, null)
        ops.cfg_euler_step(xe, v.contiguous(), 5.0, dt)
        eager_states.append(xe.clone())
    g, sx, st = m._euler_step_graph(noise.clone(), th, tp, null, 5.0, dt)
    print("static_x == noise before first replay:", float((sx - noise).abs().max()))
    for i, t in enumerate(ts):
        st.copy_(t.repeat(2 * B)); g.replay(); torch.cuda.synchronize()
        print(f"step {i}: |x_graph - x_eager| = {float((sx - eager_states[i]).abs().max()):.3e}  (|x| max {float(eager_states[i].abs().max()):.2f})")
    # second use of the cached capture
    g2, sx2, st2 = m._euler_step_graph(noise.clone(), th, tp, null, 5.0, dt)
    print("cache hit:", g2 is g)
    for i, t in enumerate(ts):
        st2.copy_(t.repeat(2 * B)); g2.replay()
    torch.cuda.synchronize()
    print("second run final diff", float((sx2 - eager_states[-1]).abs().max()))
    out_e = None
import src.models.diff_model as DM
DM.SAMPLE_GRAPH = False
a = m.sample_imgs(B, steps, "a prompt", cfg_scale=5.0, width=8 * L, height=8 * L, generator=torch.Generator().manual_seed(8))
DM.SAMPLE_GRAPH = True
b = m.sample_imgs(B, steps, "a prompt", cfg_scale=5.0, width=8 * L, height=8 * L, generator=torch.Generator().manual_seed(8))
print("sample_imgs eager vs graph:", float((a - b).abs().max()))

"""Debug: eager forward vs captured-and-replayed forward under no_grad (per-block outputs)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
m = diff_model(device=dev, **cfg).eval()
if len(sys.argv) > 1 and sys.argv[1] == "synth":
    from oracle import mmdit_oracle as O
    m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
B = 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
x = torch.randn(B, 16, L, L, device=dev)
t = torch.full((B,), 0.7, device=dev)
c = torch.randn(B, 154, 2304, device=dev).half()
pooled = torch.randn(B, 768, device=dev).half()
null = torch.tensor([0, 0, 1, 1]).bool().to(dev)
rec = []
hooks = [blk.register_forward_hook(lambda mod, inp, out: rec.append((out[0], out[1]))) for blk in m.blocks]
with torch.no_grad():
    v0 = m(x, t, c.clone(), pooled.clone(), null, null, null).float().clone()
    eager = [(a.float().clone(), b.float().clone()) for a, b in rec]; rec.clear()
    v1 = m(x, t, c.clone(), pooled.clone(), null, null, null).float().clone()
    print("eager vs eager", float((v0 - v1).abs().max())); rec.clear()
    sx, st, sc, sp = x.clone(), t.clone(), c.clone(), pooled.clone()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m(sx, st, sc, sp, null, null, null)
    torch.cuda.current_stream().wait_stream(s); rec.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        vg = m(sx, st, sc, sp, null, null, null)
    cap = list(rec)
    for it in range(2):
        g.replay(); torch.cuda.synchronize()
        print(f"replay {it}: v diff", float((vg.float() - v0).abs().max()), "ref max", float(v0.abs().max()))
        for i, ((a, b), (ea, eb)) in enumerate(zip(cap, eager)):
            print(f"   block {i}: x diff {float((a.float() - ea).abs().max()):.3e}  c diff {float((b.float() - eb).abs().max()):.3e}")
    st.fill_(0.3); g.replay(); torch.cuda.synchronize()
    v3 = m(x, torch.full((B,), 0.3, device=dev), c.clone(), pooled.clone(), null, null, null).float()
    print("t=0.3 replay vs eager", float((vg.float() - v3).abs().max()))

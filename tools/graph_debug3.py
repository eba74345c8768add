"""Debug: which part of the captured Euler step breaks (repeat / forward / cfg_euler)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from mmdit import ops
from oracle import mmdit_oracle as O
dev = torch.device("cuda")
torch.manual_seed(0)
dim, heads, depth, L = 256, 4, 2, 32
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
t = torch.ones(2 * B, device=dev)
rec = []
hooks = [blk.register_forward_hook(lambda mod, inp, out: rec.append((out[0], out[1]))) for blk in m.blocks]


def capture(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); rec.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out, list(rec)


with torch.no_grad():
    v_ref = m.forward(noise.repeat(2, 1, 1, 1), t, th, tp, null, null, null).float().clone()
    eager = [(a.float().clone(), b.float().clone()) for a, b in rec]; rec.clear()
    # A: repeat + forward
    sx = noise.clone()
    gA, vA, capA = capture(lambda: m.forward(sx.repeat(2, 1, 1, 1), t, th, tp, null, null, null))
    sx.copy_(noise); gA.replay(); torch.cuda.synchronize()
    print("A repeat+forward      : v diff", float((vA.float() - v_ref).abs().max()))
    for i, ((a, b), (ea, eb)) in enumerate(zip(capA, eager)):
        print(f"     block {i}: x diff {float((a.float() - ea).abs().max()):.3e}  c diff {float((b.float() - eb).abs().max()):.3e}")
    # B: forward on a pre-repeated input
    sx4 = noise.repeat(2, 1, 1, 1).contiguous()
    gB, vB, capB = capture(lambda: m.forward(sx4, t, th, tp, null, null, null))
    gB.replay(); torch.cuda.synchronize()
    print("B forward(pre-repeated): v diff", float((vB.float() - v_ref).abs().max()))
    for i, ((a, b), (ea, eb)) in enumerate(zip(capB, eager)):
        print(f"     block {i}: x diff {float((a.float() - ea).abs().max()):.3e}  c diff {float((b.float() - eb).abs().max()):.3e}")
    # C: random distinct samples (as graph_debug.py)
    xr = torch.randn(2 * B, 16, L, L, device=dev)
    v_ref2 = m.forward(xr, t, th, tp, null, null, null).float().clone(); rec.clear()
    gC, vC, _ = capture(lambda: m.forward(xr, t, th, tp, null, null, null))
    gC.replay(); torch.cuda.synchronize()
    print("C forward(random x)    : v diff", float((vC.float() - v_ref2).abs().max()))
    # D: eager twice (determinism of the eager path itself on this input)
    v_ref3 = m.forward(noise.repeat(2, 1, 1, 1), t, th, tp, null, null, null).float()
    print("D eager vs eager       : v diff", float((v_ref3 - v_ref).abs().max()))

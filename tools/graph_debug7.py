"""Debug: does any kernel read memory it did not write?  torch.empty / empty_like are poisoned with NaN
(floating dtypes) and the no_grad forward + the train step are compared with the clean runs."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from oracle import mmdit_oracle as O
dev = torch.device("cuda")
torch.manual_seed(0)
dim, heads, depth, L = (int(a) for a in (sys.argv[1:5] or [256, 4, 2, 32]))
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
_empty, _empty_like = torch.empty, torch.empty_like
poison = {"on": False, "val": float("nan")}


def p_empty(*a, **k):
    t = _empty(*a, **k)
    if poison["on"] and t.is_cuda and t.dtype.is_floating_point:
        t.fill_(poison["val"])
    return t


def p_empty_like(*a, **k):
    t = _empty_like(*a, **k)
    if poison["on"] and t.is_cuda and t.dtype.is_floating_point:
        t.fill_(poison["val"])
    return t


torch.empty, torch.empty_like = p_empty, p_empty_like
with torch.no_grad():
    v0 = m.forward(noise.repeat(2, 1, 1, 1), t1, th, tp, null, null, null).float().clone()
    for val in (float("nan"), 1e30, -7.0):
        poison.update(on=True, val=val)
        v1 = m.forward(noise.repeat(2, 1, 1, 1), t1, th, tp, null, null, null).float().clone()
        poison["on"] = False
        print(f"no_grad forward, poison {val}: diff {float((v1 - v0).abs().max())}  nan {int(torch.isnan(v1).sum())}")
# training step (grad mode): loss + a gradient
from mmdit.functional import rf_loss
m.train()
x0 = noise.repeat(2, 1, 1, 1).bfloat16()


def train_once():
    for p in m.parameters():
        p.grad = None
    torch.manual_seed(5)
    v = m(x0.float(), t1 * 0.5, th.clone(), tp.clone(), null, null, null)
    loss = rf_loss(v, torch.ones_like(x0), x0)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), m.blocks[0].MLP_x.MLP.w12.weight.grad.clone(), m.blocks[1].attn.query_proj_x.weight.grad.clone()


l0, g0, h0 = train_once()
for val in (float("nan"), 1e30):
    poison.update(on=True, val=val)
    l1, g1, h1 = train_once()
    poison["on"] = False
    print(f"train step, poison {val}: loss diff {abs(l1 - l0)}  grad diff {float((g1 - g0).abs().max())} / {float((h1 - h0).abs().max())}"
          f"  nan {int(torch.isnan(g1).sum())}")

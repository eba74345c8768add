"""Debug: one no_grad forward (no hooks) -- run under compute-sanitizer with PYTORCH_NO_CUDA_MEMORY_CACHING=1
to catch a kernel touching memory of a tensor Python has already released."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
from oracle import mmdit_oracle as O
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
m = diff_model(device=dev, **cfg).eval()
m.load_state_dict(O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
m.load_text_encoders()
B, L = 2, 32
noise = torch.randn(B, 16, L, L).to(dev).float().contiguous()
th, tp = m.text_encoders.text_to_embedding("a prompt")
null = torch.tensor([0] * B + [1] * B).bool().to(dev)
th = th.repeat(2 * B, 1, 1).to(dev); tp = tp.repeat(2 * B, 1).to(dev)
t1 = torch.ones(2 * B, device=dev)
with torch.no_grad():
    for i in range(2):
        v = m.forward(noise.repeat(2, 1, 1, 1), t1, th, tp, null, null, null)
        torch.cuda.synchronize()
        print("forward", i, "ok; |v| max", float(v.float().abs().max()))

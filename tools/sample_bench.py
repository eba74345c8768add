"""Euler/CFG sampling throughput of the product path (BASELINE configs[4] shape family):
python tools/sample_bench.py [depth dim heads latent batch steps]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from src.models.diff_model import diff_model
depth, dim, heads, latent, batch, steps = (int(x) for x in (sys.argv[1:7] or [24, 1536, 24, 64, 16, 50]))
dev = torch.device("cuda")
torch.manual_seed(0)
m = diff_model(inCh=16, class_dim=768, patch_size=2, dim=dim, hidden_scale=4.0, num_heads=heads,
               attn_type="softmax_flash", MLP_type="swiglu", num_blocks=depth, device=dev,
               positional_encoding="RoPE2d")
m.load_text_encoders()
g = torch.Generator().manual_seed(1)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = m.sample_imgs(batch, steps, "a prompt", cfg_scale=5.0, width=latent * 8, height=latent * 8,
                        sampler="euler", generator=g)
    torch.cuda.synchronize(); dt = time.time() - t0
    N = (latent // 2) ** 2
    T = N + 154
    fwd_flops = depth * (32 * T * dim * dim + 4 * T * T * dim) * 2 * batch * steps
    print(f"run {it}: {batch} images, {steps} Euler steps (CFG, forward batch {2*batch}, T={T}): {dt:.3f} s "
          f"-> {batch/dt:.2f} img/s, ~{fwd_flops/dt/1e12:.0f} TFLOP/s, finite={bool(torch.isfinite(out).all())}")

"""GPU diagnostic: product diff_model vs the fp32 oracle on identical weights / batch.
python tools/model_probe.py [cfg1|ragged|mid]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from oracle import mmdit_oracle as O  # noqa: E402
from src.models.diff_model import diff_model  # noqa: E402
from mmdit.functional import rf_loss  # noqa: E402

CFGS = {
    "cfg1": dict(model=dict(inCh=4, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
                            attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2,
                            positional_encoding="RoPE2d"), B=2, h=32, w=32, M=154),
    "ragged": dict(model=dict(inCh=16, class_dim=768, patch_size=2, dim=128, hidden_scale=4.0, num_heads=2,
                              attn_type="softmax_flash", MLP_type="swiglu", num_blocks=3,
                              positional_encoding="RoPE2d"), B=3, h=24, w=40, M=154),
    "mid": dict(model=dict(inCh=16, class_dim=768, patch_size=2, dim=768, hidden_scale=4.0, num_heads=12,
                           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=3,
                           positional_encoding="RoPE2d"), B=4, h=32, w=32, M=154),
}


def main(name):
    cfg = CFGS[name]
    dev = torch.device("cuda")
    model = diff_model(device=dev, **cfg["model"])
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.synth_state_dict(shapes)
    model.load_state_dict(sd, strict=True)
    batch = O.synth_batch(cfg["B"], cfg["model"]["inCh"], cfg["h"], cfg["w"], cfg["M"], seed=1000)
    bd = {k: v.to(dev) for k, v in batch.items()}
    t = bd["t"]
    x_t = (1 - t)[:, None, None, None] * bd["x0"] + t[:, None, None, None] * bd["eps"]
    t0 = time.time()
    v = model(x_t, t, bd["c"].bfloat16(), bd["pooled"].bfloat16(), bd["null_pooled"], bd["null_gemma"],
              bd["null_bert"])
    loss = rf_loss(v, bd["eps"], bd["x0"])
    loss.backward()
    torch.cuda.synchronize()
    print(f"product fwd+bwd ok in {time.time() - t0:.2f}s, loss {float(loss):.6f}")

    # fp32 oracle on the GPU (plain torch), same weights and inputs
    ocfg = dict(cfg["model"], attn_type="softmax")
    P = {k: s.to(dev).requires_grad_(not k.endswith("freqs")) for k, s in sd.items()}
    lo, vo = O.rf_loss(P, ocfg, bd)
    lo.backward()
    print(f"oracle loss {float(lo):.6f}  |dloss| {abs(float(lo) - float(loss)):.2e}")
    rel = float((v.float() - vo).abs().max() / vo.abs().max())
    print(f"v_pred max-rel err {rel:.3e}")
    worst = []
    for k, p in model.named_parameters():
        go = P[k].grad
        if p.grad is None:
            print("  NO GRAD:", k, "(requires_grad", p.requires_grad, ")")
            continue
        if go is None:
            continue
        den = float(go.abs().max())
        err = float((p.grad - go).abs().max())
        worst.append((err / (den + 1e-30), k, den))
    worst.sort(reverse=True)
    for r, k, den in worst[:12]:
        print(f"  grad rel {r:.3e}  {k}  (ref max {den:.3e})")
    print("median grad rel", sorted(w[0] for w in worst)[len(worst) // 2])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cfg1")

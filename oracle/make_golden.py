"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, through oracle/ref_shim.py) on CPU with name-keyed synthetic weights and
seeded synthetic batches (oracle/mmdit_oracle.py: synth_state_dict / synth_batch).

    python oracle/make_golden.py          # only works in the build container

Each fixture holds: the ctor config, the state_dict schema (key -> shape), the reference's
velocity prediction, loss and every parameter's gradient norm (plus a few small gradients in
full) in fp32, the same under CPU bf16 autocast (the reference's own bf16 noise floor), a
short loss trajectory of the restated train step (model_trainer.py:378-503) and a fixed-seed
Euler/CFG sample (diff_model.py:367-430).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mmdit_oracle as O  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0]: tiny MMDiT (depth 2, dim 256, 4 heads), 32x32x4 latent, batch 2
    "cfg1": dict(model=dict(inCh=4, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
                            attn_type="softmax", MLP_type="swiglu", num_blocks=2, device="cpu",
                            positional_encoding="RoPE2d"), B=2, h=32, w=32, M=154),
    # non-square latent (12x20 tokens), 3 blocks, odd batch: ragged tiles everywhere
    "ragged": dict(model=dict(inCh=16, class_dim=768, patch_size=2, dim=128, hidden_scale=4.0, num_heads=2,
                              attn_type="softmax", MLP_type="swiglu", num_blocks=3, device="cpu",
                              positional_encoding="RoPE2d"), B=3, h=24, w=40, M=154),
}
# lighter fixtures (loss, output, per-gradient bf16 floors only) for two more BASELINE shapes
LIGHT = {
    # width of BASELINE configs[2]/[3] (dim 1536, 24 heads) at depth 2
    "wide": dict(model=dict(inCh=16, class_dim=768, patch_size=2, dim=1536, hidden_scale=4.0, num_heads=24,
                            attn_type="softmax", MLP_type="swiglu", num_blocks=2, device="cpu",
                            positional_encoding="RoPE2d"), B=2, h=32, w=32, M=154),
    # 512 px: 64x64 latent -> 1024 image tokens, 256 text tokens (BASELINE configs[3] wording)
    "px512": dict(model=dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
                             attn_type="softmax", MLP_type="swiglu", num_blocks=2, device="cpu",
                             positional_encoding="RoPE2d"), B=1, h=64, w=64, M=256),
}
FULL_GRADS = ["time_scale", "learnable_scalar", "learnable_scalar2", "out_proj.bias",
              "blocks.0.attn.q_norm_x.weight", "blocks.0.attn.k_norm_c.weight", "blocks.1.y_proj.0.bias",
              "blocks.0.MLP_x.MLP.w3.bias"]


def ref_step(model, batch, autocast):
    """model_trainer.py:394-446 on the reference model (noise_batch restated to reuse the stored eps)."""
    for p in model.parameters():
        p.grad = None
    t = batch["t"]
    x_t = (1 - t)[:, None, None, None] * batch["x0"] + t[:, None, None, None] * batch["eps"]
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        v = model(x_t.detach(), t, batch["c"].clone(), batch["pooled"].clone(), batch["null_pooled"],
                  batch["null_gemma"], batch["null_bert"])
        loss = F.mse_loss(v.float(), batch["eps"] - batch["x0"], reduction="none").flatten(1, -1).mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return v.detach().float(), float(loss), grads


def main():
    torch.set_num_threads(8)
    diff_model = import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, cfg in LIGHT.items():
        model = diff_model(**dict(cfg["model"], checkpoint_MLP=False, checkpoint_attn=False))
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(O.synth_state_dict(shapes), strict=True)
        batch = O.synth_batch(cfg["B"], cfg["model"]["inCh"], cfg["h"], cfg["w"], cfg["M"], seed=1000)
        v32, l32, g32 = ref_step(model, batch, False)
        v16, l16, g16 = ref_step(model, batch, True)
        out = dict(config=cfg, loss_fp32=l32, loss_bf16=l16,
                   v_absmax=float(v32.abs().max()), v_relerr_bf16=float((v16 - v32).abs().max() / v32.abs().max()),
                   gradnorm_fp32={k: float(g.norm()) for k, g in g32.items()},
                   grad_l2err_bf16={k: float((g16[k].float() - g).norm() / (g.norm() + 1e-30)) for k, g in g32.items()})
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        worst = max(out["grad_l2err_bf16"].items(), key=lambda kv: kv[1])
        print(name, "loss fp32", l32, "bf16", l16, "worst bf16 grad L2 err", worst, "->", path,
              os.path.getsize(path) // 1024, "KiB")
    for name, cfg in CONFIGS.items():
        mk = dict(cfg["model"], checkpoint_MLP=False, checkpoint_attn=False)
        model = diff_model(**mk)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = O.synth_state_dict(shapes)
        model.load_state_dict(sd, strict=True)
        C = cfg["model"]["inCh"]
        batch = O.synth_batch(cfg["B"], C, cfg["h"], cfg["w"], cfg["M"], seed=1000)
        out = dict(config=cfg, shapes=shapes)
        gsave = {}
        for tag, ac in (("fp32", False), ("bf16", True)):
            v, loss, grads = ref_step(model, batch, ac)
            gsave[tag] = grads
            out[f"v_{tag}"] = v
            out[f"loss_{tag}"] = loss
            out[f"gradnorm_{tag}"] = {k: float(g.norm()) for k, g in grads.items()}
            out[f"grads_{tag}"] = {k: grads[k] for k in FULL_GRADS if k in grads}
        # the reference's own bf16-autocast error per parameter gradient (max-abs err / max-abs fp32)
        out["grad_relerr_bf16"] = {
            k: float((gsave["bf16"][k].float() - g32).abs().max() / (g32.abs().max() + 1e-30))
            for k, g32 in gsave["fp32"].items() if k in gsave["bf16"]}
        # loss trajectory of the restated train step (fp32, AdamW 1e-4, clip 1.0), 8 steps
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
        traj = []
        for s in range(8):
            b = O.synth_batch(cfg["B"], C, cfg["h"], cfg["w"], cfg["M"], seed=2000 + s)
            _, loss, _ = ref_step(model, b, False)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            traj.append(loss)
        out["loss_traj_fp32"] = traj
        # fixed-seed Euler/CFG sample through the reference's own sample_imgs with a stub encoder
        model.load_state_dict(sd, strict=True)
        from types import SimpleNamespace
        g = torch.Generator().manual_seed(7)
        th = torch.randn((1, cfg["M"], O.TEXT_DIM), generator=g)
        tp = torch.randn((1, 768), generator=g)
        model.text_encoders = SimpleNamespace(
            VAE=SimpleNamespace(config=SimpleNamespace(latent_channels=C, shift_factor=0.0, scaling_factor=1.0),
                                dtype=torch.float32, decode=lambda z: SimpleNamespace(sample=z)),
            text_to_embedding=lambda text: (th.clone(), tp.clone()))
        gen = torch.Generator().manual_seed(11)
        # sample_imgs clamps the "decoded" latents to [-1,1]; keep the raw loop result instead by
        # scaling: store both the clamped API output and an un-clamped oracle replay input set
        smp = model.sample_imgs(2, 4, "a prompt", cfg_scale=5.0, width=cfg["h"] * 8, height=cfg["w"] * 8,
                                sampler="euler", generator=gen)
        # inputs are re-derivable: text from Generator(7) (hidden then pooled), noise from Generator(11)
        out["sample_seeds"] = dict(text=7, noise=11, steps=4, cfg_scale=5.0, batch=2)
        out["sample_euler_clamped"] = smp
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(name, "loss fp32", out["loss_fp32"], "bf16", out["loss_bf16"], "traj", [round(x, 4) for x in traj],
              "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

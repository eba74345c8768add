"""GPU (-m gpu): the SURVEY 8f "next" rows on the device -- device-resident EMA (model_trainer.py:256,
537-541), the loader wire format through HostFeed.submit_wire (model_trainer.py:353-370), the other
samplers (diff_model.py:431-460) and the captured Euler step (diff_model.py:407-430)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CFG = dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")


@pytest.fixture(scope="module")
def dev():
    from mmdit import _lib
    _lib.check(_lib.lib().mmdit_device_check(), "mmdit_device_check")
    return torch.device("cuda")


def _model(dev, seed=0):
    from oracle import mmdit_oracle as O
    from src.models.diff_model import diff_model
    m = diff_model(device=dev, **CFG)
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, salt=seed)
    m.load_state_dict(sd, strict=True)
    return m, sd


def test_device_ema_matches_the_reference_blend_and_feeds_the_kernels(dev):
    """f2: EMA blended on the device equals the reference's CPU loop bit for bit; after training with
    the fused optimizer (managed bf16 shadows) `copy_to(model)` must make the kernels see the EMA
    weights -- a model freshly loaded from the EMA state_dict gives the identical output."""
    from mmdit.ema import DeviceEMA
    from mmdit.train import RFTrainer, host_batch
    from src.models.diff_model import diff_model
    model, _ = _model(dev)
    ema = DeviceEMA(model, decay=0.9, update_freq=1)
    ref = [p.detach().cpu().clone() for p in ema.params]             # model_trainer.py:256
    tr = RFTrainer(model, lr=1e-3, ema=ema)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s in range(3):
        tr.step(tr.to_device(host_batch(2, 16, 32, 32, 154, seed=60 + s)))
        for r, p in zip(ref, ema.params):                            # :537-541 verbatim, on the CPU
            r.mul_(0.9).add_(p.detach().cpu(), alpha=1.0 - 0.9)
    for r, e in zip(ref, ema.ema):
        assert torch.equal(r, e.cpu())
    e0.record(); ema.update(); e1.record(); torch.cuda.synchronize()
    nbytes = sum(p.numel() for p in ema.params) * 4 * 3              # read ema + p, write ema
    print(f"DeviceEMA.update: {e0.elapsed_time(e1) * 1e3:.0f} us for {nbytes / 1e6:.1f} MB "
          f"({nbytes / e0.elapsed_time(e1) / 1e6:.0f} GB/s at this tiny size)")
    sd_ema = ema.state_dict()
    assert set(sd_ema) == set(model.state_dict())
    hb = host_batch(2, 16, 32, 32, 154, seed=99)
    b = tr.to_device(hb)
    x_t = torch.randn(2, 16, 32, 32, device=dev)
    with torch.no_grad():
        before = model(x_t, b["t"], b["c"].clone(), b["pooled"].clone()).float()
        ema.copy_to(model)                                           # in-place load_state_dict
        after = model(x_t, b["t"], b["c"].clone(), b["pooled"].clone()).float()
        fresh = diff_model(device=dev, **CFG)
        fresh.load_state_dict(sd_ema, strict=True)
        want = fresh(x_t, b["t"], b["c"].clone(), b["pooled"].clone()).float()
    assert torch.equal(after, want)                 # no stale bf16 shadow anywhere
    assert not torch.equal(before, after)


def test_wire_format_batch_trains_through_hostfeed(dev):
    """f3: a batch in the loader's wire format (+inf padded bf16 latents of a 24x40 bucket inside a 64x64
    frame) -> HostFeed.submit_wire -> take: the device batch equals the reference receiver's gather
    (model_trainer.py:363-370) and a trainer step runs on it (graph keyed by the bucket's geometry)."""
    from mmdit import feed
    from mmdit.train import HostFeed, RFTrainer
    model, _ = _model(dev)
    g = torch.Generator().manual_seed(5)
    B, h, w = 4, 24, 40
    x0 = torch.randn((B, 16, h, w), generator=g).to(torch.bfloat16)
    wire = {"images": feed.pad_latents(x0, 64), "text": torch.randn((B, 154, 2304), generator=g).to(torch.bfloat16),
            "text_pooled": torch.randn((B, 768), generator=g).to(torch.bfloat16)}
    hf = HostFeed(dev)
    torch.manual_seed(3)
    hf.submit_wire(wire)
    batch = hf.take()
    on_dev = wire["images"].to(dev)
    ref = on_dev[on_dev != torch.inf].reshape(B, 16, h, w)           # the reference's mask gather, on the device
    assert torch.equal(batch["x0"], ref) and batch["x0"].is_contiguous()
    assert torch.equal(batch["c"].cpu(), wire["text"]) and torch.equal(batch["pooled"].cpu(), wire["text_pooled"])
    tr = RFTrainer(model, use_graph=True)
    l1 = float(tr.step(batch))
    torch.manual_seed(4)
    hf.submit_wire({**wire, "images": feed.pad_latents(x0[:, :, :16, :16].contiguous(), 64)})   # another bucket
    l2 = float(tr.step(hf.take()))
    assert l1 == l1 and l2 == l2 and len(tr._graphs) == 2


@pytest.mark.parametrize("sampler", ["heun", "euler_stochastic"])
def test_other_samplers_on_the_device(dev, sampler):
    """f4: Heun against the oracle's Euler machinery restated for Heun (diff_model.py:446-460), and the
    stochastic Euler sampler finite and seeded-reproducible."""
    from oracle import mmdit_oracle as O
    model, sd = _model(dev)
    model.load_text_encoders()
    gen = lambda: torch.Generator().manual_seed(21)
    out = model.sample_imgs(2, 6, "a prompt", cfg_scale=3.0, width=128, height=128, sampler=sampler, generator=gen())
    assert out.shape == (2, 16, 16, 16) and bool(torch.isfinite(out).all())
    again = model.sample_imgs(2, 6, "a prompt", cfg_scale=3.0, width=128, height=128, sampler=sampler, generator=gen())
    assert float((out - again).abs().max()) <= 2e-2           # same seed -> same sample (fp32 atomics aside)
    if sampler != "heun":
        return
    th, tp = model.text_encoders.text_to_embedding("a prompt")
    P = {k: v.to(dev) for k, v in sd.items()}
    cfg = dict(CFG, attn_type="softmax")
    x = torch.randn((2, 16, 16, 16), generator=gen()).to(dev).float()
    null = torch.tensor([0, 0, 1, 1]).bool().to(dev)
    thr, tpr = th.float().repeat(4, 1, 1).to(dev), tp.float().repeat(4, 1).to(dev)

    def vel(xx, t):
        v = O.forward(P, cfg, xx.repeat(2, 1, 1, 1), t.repeat(4).to(dev), thr, tpr, null, null, null)
        return (1 + 3.0) * v[:2] - 3.0 * v[2:]
    dt = 1 / 6
    with torch.no_grad():
        for t in torch.linspace(1, dt, 6):
            v1 = vel(x, t)
            v2 = vel(x - v1 * dt, t - dt)
            x = x - (v1 + v2) * (dt / 2)
    ref = x.clamp(-1, 1)
    mse = float(((out - ref) ** 2).mean())
    psnr = float(10 * torch.log10(torch.tensor(4.0 / max(mse, 1e-12))))
    assert psnr >= 30.0, psnr


def test_captured_euler_step_equals_the_eager_loop(dev, monkeypatch):
    """The sampler replays ONE captured Euler step (2B forward + fused CFG update); the result must equal
    the eager loop, and a weight reload between two calls must invalidate the capture."""
    import src.models.diff_model as DM
    model, sd = _model(dev)
    model.load_text_encoders()
    run = lambda: model.sample_imgs(2, 5, "a prompt", cfg_scale=5.0, width=128, height=128, sampler="euler",
                                    generator=torch.Generator().manual_seed(8))
    monkeypatch.setattr(DM, "SAMPLE_GRAPH", False)
    eager = run()
    monkeypatch.setattr(DM, "SAMPLE_GRAPH", True)
    graphed = run()
    assert float((eager - graphed).abs().max()) <= 1e-3
    assert len(model.__dict__["_sample_graphs"]) == 1
    graphed2 = run()                                   # replay of the cached capture
    assert float((graphed - graphed2).abs().max()) <= 1e-3
    other, sd2 = _model(dev, seed=1)
    model.load_state_dict(sd2)                         # in place: versions move, capture must be redone
    other.load_text_encoders()
    monkeypatch.setattr(DM, "SAMPLE_GRAPH", False)
    want = other.sample_imgs(2, 5, "a prompt", cfg_scale=5.0, width=128, height=128, sampler="euler",
                             generator=torch.Generator().manual_seed(8))
    monkeypatch.setattr(DM, "SAMPLE_GRAPH", True)
    got = run()
    assert float((want - got).abs().max()) <= 1e-3
    import copy
    assert "_sample_graphs" in model.__dict__ and len(copy.deepcopy(model).__dict__["_sample_graphs"]) == 0


def test_no_kernel_reads_memory_it_did_not_write(dev, monkeypatch):
    """Every buffer the product allocates with torch.empty / empty_like is poisoned (NaN, then 1e30): the
    no_grad forward must be bit-identical and a train step's loss / gradients unchanged up to fp32 atomics
    order -- partial tiles, ragged rows and workspaces never leak uninitialised memory into a result."""
    from mmdit.functional import rf_loss
    model, _ = _model(dev)
    B, L = 3, 24
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 16, L, 40, generator=g).to(dev)                 # ragged 12x20 token grid
    c = torch.randn(B, 154, 2304, generator=g).to(dev).bfloat16()
    pooled = torch.randn(B, 768, generator=g).to(dev).bfloat16()
    t = torch.rand(B, generator=g).to(dev)
    _empty, _empty_like = torch.empty, torch.empty_like
    state = {"val": None}

    def fill(tn):
        if state["val"] is not None and tn.is_cuda and tn.dtype.is_floating_point:
            tn.fill_(state["val"])
        return tn
    monkeypatch.setattr(torch, "empty", lambda *a, **k: fill(_empty(*a, **k)))
    monkeypatch.setattr(torch, "empty_like", lambda *a, **k: fill(_empty_like(*a, **k)))

    def run_train():
        for p in model.parameters():
            p.grad = None
        v = model(x, t, c.clone(), pooled.clone())
        loss = rf_loss(v, torch.ones_like(x), x)
        loss.backward()
        return float(loss), model.blocks[0].MLP_c.MLP.w3.weight.grad.clone(), model.blocks[1].attn.key_proj_x.weight.grad.clone()

    with torch.no_grad():
        v0 = model(x, t, c.clone(), pooled.clone()).float().clone()
    l0, g0, h0 = run_train()
    for val in (float("nan"), 1e30):
        state["val"] = val
        with torch.no_grad():
            v1 = model(x, t, c.clone(), pooled.clone()).float().clone()
        l1, g1, h1 = run_train()
        state["val"] = None
        assert torch.equal(v0, v1), val
        assert abs(l1 - l0) <= 1e-5 and bool(torch.isfinite(g1).all())
        assert float((g1 - g0).abs().max()) <= 1e-5 * float(g0.abs().max()) + 1e-9
        assert float((h1 - h0).abs().max()) <= 1e-5 * float(h0.abs().max()) + 1e-9

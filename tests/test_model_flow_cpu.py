"""CPU: the product's HOST logic end to end -- reference-shaped modules, autograd glue, packed
weights, gradient buckets -- executed with tests/cpu_ops.py standing in for the CUDA kernels and
compared with the fp32 oracle and the reference's golden outputs at the GPU parity tolerances.
What this pins without a GPU: the order and wiring of every op in forward and backward (residuals,
modulation slots, stream splits, bias / norm-weight gradients), i.e. everything except the kernels
themselves (those are checked op by op against the same torch formulas in the -m gpu tests)."""
import contextlib

import pytest
import torch

import cpu_ops
from mmdit import functional, ops, streams
from oracle import mmdit_oracle as O
from src.models.diff_model import diff_model


@pytest.fixture
def cpu_kernels(monkeypatch):
    for name in cpu_ops.ALL:
        if hasattr(ops, name):
            monkeypatch.setattr(ops, name, getattr(cpu_ops, name))
    monkeypatch.setattr(ops, "_gemm", cpu_ops.gemm)
    yield


def _run(cfg, golden_entry, seed=1000):
    m = cfg["model"]
    model = diff_model(**dict(m, attn_type="softmax_flash", device="cpu"))
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd, strict=True)
    b = O.synth_batch(cfg["B"], m["inCh"], cfg["h"], cfg["w"], cfg["M"], seed=seed)
    t = b["t"]
    x_t = (1 - t)[:, None, None, None] * b["x0"] + t[:, None, None, None] * b["eps"]
    v = model(x_t, t, b["c"].bfloat16(), b["pooled"].bfloat16(), b["null_pooled"], b["null_gemma"], b["null_bert"])
    loss = functional.rf_loss(v, b["eps"], b["x0"])
    loss.backward()
    return model, v, float(loss)


def _check_against_golden(model, v, loss, g):
    assert abs(loss - g["loss_fp32"]) <= 1e-3
    ref_v = g["v_fp32"]
    assert float((v.float() - ref_v).abs().max() / ref_v.abs().max()) <= 2e-2
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, k
        n_ref = g["gradnorm_fp32"][k]
        floor = g["grad_relerr_bf16"].get(k, 0.0)
        if n_ref == 0.0:
            assert float(p.grad.abs().max()) == 0.0, k
            continue
        n = float(p.grad.float().norm())
        # (6e-2 rather than the GPU tests' 4e-2: the stand-in ops round to bf16 at slightly different
        # points than the kernels, which the cancellation-heavy scalar gradients feel)
        assert abs(n - n_ref) <= max(6e-2, 3 * floor) * n_ref + 1e-7, (k, n, n_ref)


@pytest.mark.parametrize("name", ["cfg1", "ragged"])
def test_modules_wire_the_ops_like_the_reference(cpu_kernels, golden, name):
    g = golden(name)
    model, v, loss = _run(g["config"], g)
    _check_against_golden(model, v, loss, g)


def test_unfused_swiglu_and_fused_qknorm_paths_agree(cpu_kernels, golden, monkeypatch):
    """The optional schedules are the same function: SwiGLU activation outside the GEMM epilogue,
    and (experimental) QK-norm + RoPE inside the q|k|v projection's epilogue, which only exists on
    the two-stream path -- driven here through stand-in streams."""
    g = golden("cfg1")
    base_model, base_v, base_loss = _run(g["config"], g)

    # gated residual + LayerNorm-modulate as two kernels instead of the fused pass: same function
    monkeypatch.setattr(functional, "FUSED_GATE_LN", False)
    m1, v1, loss1 = _run(g["config"], g)
    assert torch.equal(v1, base_v) and loss1 == base_loss
    for (k, p), (_, q) in zip(m1.named_parameters(), base_model.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p.grad, q.grad), k
    monkeypatch.setattr(functional, "FUSED_GATE_LN", True)

    # SwiGLU backward inside w3's data-gradient GEMM (the activation gradient only exists as a dummy view
    # between the two autograd nodes): same function, bit for bit; shape gate lifted for the tiny model
    _ops = functional.ops            # the stand-in module the fixture installed
    monkeypatch.setattr(functional, "FUSED_SWIGLU_BWD", True)
    monkeypatch.setattr(_ops, "swiglu_bwd_fusable", lambda rows, hidden: True)
    calls = []
    real = _ops.gemm_swiglu_bwd
    monkeypatch.setattr(_ops, "gemm_swiglu_bwd", lambda *a: (calls.append(1), real(*a))[1])
    m3, v3, loss3 = _run(g["config"], g)
    assert calls, "the fused path was not taken"
    assert torch.equal(v3, base_v) and loss3 == base_loss
    for (k, p), (_, q) in zip(m3.named_parameters(), base_model.named_parameters()):
        if p.requires_grad:
            assert torch.equal(p.grad, q.grad), k
    monkeypatch.setattr(functional, "FUSED_SWIGLU_BWD", False)

    monkeypatch.setattr(functional, "FUSED_SWIGLU", False)
    m2, v2, loss2 = _run(g["config"], g)
    assert abs(loss2 - base_loss) <= 2e-3
    _check_against_golden(m2, v2, loss2, g)
    monkeypatch.setattr(functional, "FUSED_SWIGLU", True)

    # two-stream schedule with inert stream objects (no CUDA here): exercises _forward_two_streams
    class _S:
        device = torch.device("cpu")

        def wait_stream(self, other):
            pass

    import src.blocks.Transformer_Block_Dual as TBD
    monkeypatch.setattr(streams, "active", lambda t: True)
    monkeypatch.setattr(streams, "side", lambda dev: _S())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _S())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    m3, v3, loss3 = _run(g["config"], g)
    assert abs(loss3 - base_loss) <= 1e-6          # same ops in the same order on one "device"
    assert torch.equal(v3, base_v)

    monkeypatch.setattr(TBD, "FUSED_QKNORM", True)
    m4, v4, loss4 = _run(g["config"], g)
    assert abs(loss4 - base_loss) <= 1e-6 and torch.equal(v4, base_v)
    for (k, p), (_, q) in zip(m4.named_parameters(), base_model.named_parameters()):
        if p.requires_grad:
            assert torch.allclose(p.grad, q.grad, rtol=0, atol=0), k


def test_trainer_buckets_take_the_gradients_in_place(cpu_kernels, golden):
    """world_size 1 buckets (no process group): after a backward every .grad IS its bucket slot, the
    packed weight gradients were written there directly, and the values equal the plain run's."""
    from mmdit.train import GradBuckets
    g = golden("cfg1")
    base_model, _, _ = _run(g["config"], g)
    m = g["config"]["model"]
    model = diff_model(**dict(m, attn_type="softmax_flash", device="cpu"))
    model.load_state_dict(base_model.state_dict())
    adjacent = [blk._mod_weights() for blk in model.blocks]
    gb = GradBuckets(list(model.named_parameters()), 1, None, torch.device("cpu"), peer=False, adjacent=adjacent)
    gb.install_hooks()
    gb.reset()
    b = O.synth_batch(g["config"]["B"], m["inCh"], g["config"]["h"], g["config"]["w"], g["config"]["M"], seed=1000)
    t = b["t"]
    x_t = (1 - t)[:, None, None, None] * b["x0"] + t[:, None, None, None] * b["eps"]
    v = model(x_t, t, b["c"].bfloat16(), b["pooled"].bfloat16(), b["null_pooled"], b["null_gemma"], b["null_bert"])
    functional.rf_loss(v, b["eps"], b["x0"]).backward()
    gb.finish()
    total = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert gb.direct_elems + gb.copied_elems == total
    # every block-level weight gradient lands in place; what is copied in is the front-end (text
    # projections: 18 % of this tiny model, < 2 % of cfg2), biases, norm weights and the batched y_proj
    assert gb.direct_elems > 0.75 * total
    for (k, p), (_, q) in zip(model.named_parameters(), base_model.named_parameters()):
        if p.requires_grad:
            assert p.grad.data_ptr() == p._grad_slot.data_ptr(), k
            assert torch.equal(p.grad, q.grad), k


@pytest.mark.parametrize("sampler", ["euler", "heun"])
def test_sampler_loop_wires_cfg_and_euler_like_the_reference(cpu_kernels, sampler):
    """sample_imgs (diff_model.py:367-480): repeat-2 batch, null conditioning for the second half,
    CFG combine + Euler update, width/height swap -- against the oracle's sampler on CPU."""
    cfg = dict(inCh=16, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
               attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
    model = diff_model(device="cpu", **cfg)
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd, strict=True)
    model.load_text_encoders()
    out = model.sample_imgs(2, 3, "a prompt", cfg_scale=5.0, width=128, height=128, sampler=sampler,
                            generator=torch.Generator().manual_seed(11))
    assert out.shape == (2, 16, 16, 16) and bool(torch.isfinite(out).all())
    if sampler != "euler":
        return
    th, tp = model.text_encoders.text_to_embedding("a prompt")
    noise = torch.randn((2, 16, 16, 16), generator=torch.Generator().manual_seed(11))
    ref = O.sample_euler(dict(sd), dict(cfg, attn_type="softmax"), noise, th, tp, 3, 5.0).clamp(-1, 1)
    mse = float(((out - ref) ** 2).mean())
    psnr = 10 * torch.log10(torch.tensor(4.0 / max(mse, 1e-12)))
    assert float(psnr) >= 30.0, float(psnr)


def test_trainer_step_order_matches_the_reference_loop(cpu_kernels):
    """RFTrainer's eager step (zero -> forward -> loss -> backward -> clip 1.0 -> AdamW), driven with the
    oracle's noise, against the oracle's restatement of model_trainer.py:463-503 over a few steps.
    (torch.optim.AdamW here: the fused clip+AdamW kernel is a -m gpu test of its own.)"""
    from mmdit.train import RFTrainer
    cfg = dict(inCh=4, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
               attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
    model = diff_model(device="cpu", **cfg)
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
    model.load_state_dict(sd, strict=True)
    tr = RFTrainer(model, fused_optimizer=False)
    oracle = O.TrainOracle(sd, dict(cfg, attn_type="softmax"))
    for s in range(4):
        b = O.synth_batch(2, 4, 32, 32, 154, seed=3000 + s)
        lo = oracle.step(b)
        tr._zero()
        x_t = ops.rf_noise(b["x0"].contiguous(), b["eps"].contiguous(), b["t"])
        v = model(x_t, b["t"], b["c"].bfloat16(), b["pooled"].bfloat16(), b["null_pooled"], b["null_gemma"],
                  b["null_bert"])
        loss = functional.rf_loss(v, b["eps"], b["x0"])
        loss.backward()
        tr._update()
        tr._after_step()
        assert abs(float(loss) - lo) <= 1e-3 * max(1.0, abs(lo)), (s, float(loss), lo)
    assert tr.steps_done == 4


def test_block_bucket_exchange_starts_before_the_next_block_backward(cpu_kernels, golden):
    """Overlap contract of the data-parallel exchange (model_trainer.py:224 -- DDP's reducer): the
    bucket of block k must be complete, and its exchange launched, before the first gradient hook
    of block k-1 fires.  The batched y projections (computed for all blocks ahead of block 0) live in
    the "rest" bucket, which closes last -- inside the block buckets they would delay every launch
    to the end of the backward."""
    from mmdit.train import GradBuckets
    g = golden("cfg1")
    m = dict(g["config"]["model"], num_blocks=3)
    model = diff_model(**dict(m, attn_type="softmax_flash", device="cpu"))
    adjacent = [blk._mod_weights() for blk in model.blocks]
    gb = GradBuckets(list(model.named_parameters()), 1, None, torch.device("cpu"), peer=False, adjacent=adjacent)
    assert all(GradBuckets.bucket_key(f"blocks.{i}.y_proj.0.{w}") == ("rest", 0) for i in range(3) for w in ("weight", "bias"))
    gb.install_hooks()
    gb.world_size = 2                       # hooks launch only when there is somebody to exchange with
    events = []
    orig_ready = gb._ready

    def launch(bi):
        events.append(("launch", gb.buckets[bi][0]))
        gb._works[bi] = True

    def ready(bi, p):
        events.append(("hook", gb.buckets[bi][0]))
        orig_ready(bi, p)

    gb._launch, gb._ready = launch, ready
    gb.reset()
    b = O.synth_batch(2, m["inCh"], 32, 32, 154, seed=1000)
    t = b["t"]
    x_t = (1 - t)[:, None, None, None] * b["x0"] + t[:, None, None, None] * b["eps"]
    v = model(x_t, t, b["c"].bfloat16(), b["pooled"].bfloat16(), b["null_pooled"], b["null_gemma"], b["null_bert"])
    functional.rf_loss(v, b["eps"], b["x0"]).backward()
    launches = [k for kind, k in events if kind == "launch"]
    assert launches[:3] == [("block", 2), ("block", 1), ("block", 0)], launches
    for blk in (2, 1):
        i_launch = events.index(("launch", ("block", blk)))
        first_next = events.index(("hook", ("block", blk - 1)))
        assert i_launch < first_next, (blk, i_launch, first_next)

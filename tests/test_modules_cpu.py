"""CPU: host-side contract of the reference-shaped modules -- constructor signatures, state_dict
schema (checked against the schema recorded from the UNMODIFIED reference in the golden files),
checkpoint round trip, deep copy / .cpu(), out-of-scope flags, and the no-CPU-fallback rule."""
import copy
import inspect
import json
import os

import pytest
import torch

from src.blocks.Attention import Attention
from src.blocks.MLP import MLP
from src.blocks.Norm import Norm
from src.blocks.Transformer_Block_Dual import Transformer_Block_Dual
from src.models.diff_model import diff_model

TINY = dict(inCh=4, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
            attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, device="cpu",
            positional_encoding="RoPE2d")


def test_state_dict_schema_equals_reference(golden):
    for name in ("cfg1", "ragged"):
        g = golden(name)
        cfg = dict(g["config"]["model"], attn_type="softmax_flash")
        m = diff_model(**cfg)
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == g["shapes"]
        assert sum(p.numel() for p in m.parameters()) == sum(int(torch.tensor(s).prod()) for s in g["shapes"].values())
        # frozen rotary frequencies are in the state_dict but not trainable
        assert not m.blocks[0].attn.rotary_emb.freqs.requires_grad


def test_constructor_signatures_match_reference():
    # parameter names in reference order (diff_model.py:83, Transformer_Block_Dual.py:15, Attention.py:16, ...)
    assert list(inspect.signature(diff_model.__init__).parameters)[1:] == [
        "inCh", "class_dim", "patch_size", "dim", "hidden_scale", "num_heads", "attn_type", "MLP_type",
        "num_blocks", "device", "positional_encoding", "max_res_orig", "max_res", "update_max_res",
        "kv_merge_attn", "qk_half_dim", "text_loss", "checkpoint_MLP", "checkpoint_attn", "start_step", "wandb_id"]
    assert list(inspect.signature(Transformer_Block_Dual.__init__).parameters)[1:] == [
        "dim", "c_dim", "hidden_scale", "num_heads", "attn_type", "MLP_type", "causal", "positional_encoding",
        "RoPE_Scale", "kv_merge_attn", "qk_half_dim", "checkpoint_MLP", "checkpoint_attn", "layer_idx", "last"]
    assert list(inspect.signature(Attention.__init__).parameters)[1:] == [
        "dim", "num_heads", "attn_type", "causal", "emb_dim", "positional_encoding", "RoPE_Scale",
        "kv_merge_attn", "qk_half_dim", "layer_idx", "dual", "last"]
    assert list(inspect.signature(MLP.__init__).parameters)[1:] == ["dim", "hidden_scale", "act"]
    assert list(inspect.signature(Norm.__init__).parameters)[1:] == ["dim", "c_dim"]
    assert list(inspect.signature(diff_model.forward).parameters)[1:] == [
        "x_t", "t", "c", "c_pooled", "nullCls_pooled", "nullCls_gemma", "nullCls_bert"]
    assert list(inspect.signature(diff_model.sample_imgs).parameters)[1:] == [
        "batchSize", "num_steps", "text_input", "cfg_scale", "width", "height", "save_intermediate",
        "use_tqdm", "sampler", "generator"]


def test_checkpoint_round_trip_and_json(tmp_path):
    m = diff_model(**TINY)
    m.wandb_id = "abc"
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    m.saveModel(str(tmp_path), EMA_state_dict=m.state_dict(), optimizer=opt, step=7)
    files = sorted(os.listdir(tmp_path))
    assert files == ["model_7s.pkl", "model_ema_7s.pkl", "model_params_7s.json", "optim_7s.pkl"]
    D = json.load(open(tmp_path / "model_params_7s.json"))
    assert D["start_step"] == 7 and D["device"] == "cpu" and D["wandb_id"] == "abc" and D["dim"] == 256
    # infer.py builds a dummy model and lets loadModel rebuild it from the JSON (infer.py:66-94)
    dummy = diff_model(**dict(TINY, dim=128, num_heads=2, num_blocks=1))
    dummy.loadModel(str(tmp_path), "model_7s.pkl", "model_params_7s.json")
    assert dummy.start_step == 7 and len(dummy.blocks) == 2
    for (k, a), (_, b) in zip(m.state_dict().items(), dummy.state_dict().items()):
        assert torch.equal(a, b), k


def test_deepcopy_cpu_and_last_block_layout():
    m = diff_model(**TINY)
    ema = copy.deepcopy(m).cpu()          # model_trainer.py:256
    assert set(ema.state_dict()) == set(m.state_dict())
    last = m.blocks[-1]
    assert last.last and not hasattr(last, "MLP_c") and not hasattr(last.attn, "out_proj_c")
    assert hasattr(m.blocks[0], "MLP_c") and hasattr(m.blocks[0].attn, "out_proj_c")
    assert m.inCh == 4 and m.class_dim == 768 and m.patch_size == 2 and m.dev == "cpu"


def test_out_of_scope_flags_raise():
    with pytest.raises(NotImplementedError):
        diff_model(**dict(TINY, attn_type="cosine"))
    with pytest.raises(NotImplementedError):
        diff_model(**dict(TINY, MLP_type="gelu"))
    with pytest.raises(NotImplementedError):
        diff_model(**dict(TINY, positional_encoding="absolute"))
    with pytest.raises(NotImplementedError):
        diff_model(**dict(TINY, kv_merge_attn=True))
    with pytest.raises(AssertionError):
        diff_model(**dict(TINY, positional_encoding="bogus"))


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    m = diff_model(**TINY)
    x = torch.randn(2, 4, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA|cuda|B200"):
        m(x, torch.rand(2), torch.randn(2, 154, 2304), torch.randn(2, 768))


def test_rotary_tables_match_oracle_angles():
    from oracle import mmdit_oracle as O
    m = diff_model(**TINY)
    rot = m.blocks[0].attn.rotary_emb
    cos, sin = rot.tables(12, 20)
    ang = O.axial_angles(rot.freqs.detach(), 12, 20).reshape(240, 64)[:, 0::2]
    assert torch.allclose(cos, ang.cos()) and torch.allclose(sin, ang.sin())
    assert rot.get_axial_freqs(12, 20).shape == (12, 20, 64)


def test_synthetic_text_encoder_stub_shapes():
    m = diff_model(**TINY)
    m.load_text_encoders()
    h, p = m.text_encoders.text_to_embedding("a photo of a cat")
    assert h.shape == (1, 154, 2304) and p.shape == (1, 768) and h.dtype == torch.float16
    h2, _ = m.text_encoders.text_to_embedding("a photo of a cat")
    assert torch.equal(h, h2)
    assert m.text_encoders.VAE.config.latent_channels == 16


def test_fused_gate_ln_width_limit_follows_the_row_kernel_generation(monkeypatch):
    """Host logic of the one-pass gated residual + LayerNorm-modulate (Transformer_Block_Dual.py:64-72):
    rows up to 1536 columns take it with the second-generation kernels, up to 1024 with the first; wider
    rows go through the two kernels."""
    import types

    import torch
    from mmdit import ops
    from src.blocks import Transformer_Block_Dual as T

    monkeypatch.setattr(ops, "_row_generation", 2)
    assert ops.fused_gate_ln_max_columns() == 1536
    monkeypatch.setattr(ops, "_row_generation", 1)
    assert ops.fused_gate_ln_max_columns() == 1024

    taken = []

    class FusedFn:
        @staticmethod
        def apply(a, wb, bb, gate, resid2, shift, scale, rpb, w, b):
            taken.append("fused")
            return resid2, resid2

    monkeypatch.setattr(T, "GatedLinearLNFn", FusedFn)
    monkeypatch.setattr(T, "packed_weight", lambda lin, tag, ws: ws[0])
    monkeypatch.setattr(T, "modulate_keep", lambda x, shift, scale: (taken.append("two kernels"), (x, x))[1])
    monkeypatch.setattr(T.Fn, "FUSED_GATE_LN", True)
    fake = types.SimpleNamespace(_gated=lambda a, lin, gate, resid, rpb: resid)
    for gen, width, want in ((2, 1536, "fused"), (1, 1536, "two kernels"), (1, 1024, "fused"), (2, 1792, "two kernels")):
        monkeypatch.setattr(ops, "_row_generation", gen)
        taken.clear()
        lin = types.SimpleNamespace(weight=torch.zeros(1), bias=None)
        resid = torch.zeros(2, 3, width)
        T.Transformer_Block_Dual._gated_ln(fake, resid, lin, None, resid, None, None, 3)
        assert taken == [want], (gen, width, taken)

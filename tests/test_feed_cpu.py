"""CPU: the loader -> trainer wire format (mmdit/feed.py) against the reference's own expressions
(sender src/helpers/VAE_T5_CLIP.py:438, receiver src/model_trainer.py:363-370), restated verbatim here."""
import pytest
import torch

from mmdit import feed

INF = float("inf")


def reference_receive(batch_x_0, batchSize, inCh):
    """model_trainer.py:363-370, verbatim (on whatever device the tensor lives)."""
    orig_shape = (
        batchSize,
        inCh,
        batch_x_0.shape[2] - (batch_x_0[0, 0] == torch.inf).sum(-2)[0].item(),
        batch_x_0.shape[3] - (batch_x_0[0, 0] == torch.inf).sum(-1)[0].item(),
    )
    return batch_x_0[batch_x_0 != torch.inf].reshape(orig_shape), orig_shape


@pytest.mark.parametrize("h,w", [(32, 32), (24, 40), (40, 24), (1, 1), (64, 8), (17, 63)])
def test_pad_infer_crop_equals_the_reference_receiver(h, w):
    g = torch.Generator().manual_seed(h * 100 + w)
    B, C, side = 3, 16, 64
    x0 = torch.randn((B, C, h, w), generator=g).to(torch.bfloat16)
    wire = feed.pad_latents(x0, side)                       # VAE_T5_CLIP.py:438
    assert wire.shape == (B, C, side, side) and wire.dtype == torch.bfloat16
    ref, ref_shape = reference_receive(wire, B, C)
    assert feed.infer_latent_shape(wire) == (ref_shape[2], ref_shape[3]) == (h, w)
    got = feed.crop_latents(wire, h, w)
    assert got.is_contiguous() and torch.equal(got, ref) and torch.equal(got, x0)


def test_from_wire_builds_a_trainer_batch_and_draws_like_the_reference():
    B, C, side = 4, 16, 32
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn((B, C, 24, 20), generator=g).to(torch.bfloat16)
    wire = {"images": feed.pad_latents(x0, side), "text": torch.randn((B, 154, 2304), generator=g).to(torch.bfloat16),
            "text_pooled": torch.randn((B, 768), generator=g).to(torch.bfloat16)}
    torch.manual_seed(11)
    hb, hw = feed.from_wire(wire, pin=False)
    assert hw == (24, 20)
    assert set(hb) == {"x0_padded", "c", "pooled", "t", "null_pooled", "null_gemma", "null_bert"}
    # same CPU draws, same order as model_trainer.py:378-387 (TimeSampler, then the three torch.rand)
    torch.manual_seed(11)
    t = torch.sigmoid(torch.randn(B))
    pp, pg, pb = torch.rand(B), torch.rand(B), torch.rand(B)
    assert torch.equal(hb["t"], t)
    assert torch.equal(hb["null_pooled"], pp < 0.1) and torch.equal(hb["null_gemma"], pg < 0.316)
    assert torch.equal(hb["null_bert"], pb < 0.316)
    dev = feed.finish_on_device(hb, hw)                      # "device" = CPU here: same slice copy
    assert "x0_padded" not in dev and torch.equal(dev["x0"], x0)


def test_bad_wire_batches_raise():
    with pytest.raises(ValueError):
        feed.pad_latents(torch.zeros(1, 4, 40, 8), 32)
    with pytest.raises(ValueError):
        feed.infer_latent_shape(torch.full((1, 4, 8, 8), INF))
    with pytest.raises(ValueError):
        feed.from_wire({"images": torch.zeros(2, 4, 8, 8), "text": torch.zeros(3, 154, 2304),
                        "text_pooled": torch.zeros(2, 768)}, pin=False)

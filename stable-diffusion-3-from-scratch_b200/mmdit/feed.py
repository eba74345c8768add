"""The data-feed boundary of the train step: the reference's loader -> trainer wire
format and its unpacking, restated so that nothing on the trainer side synchronises the device.

Reference (sender, src/helpers/VAE_T5_CLIP.py:438,447-478): per model rank a dict
    {"images": bf16 (B, C, max_res/8, max_res/8), right/bottom padded with +inf to the bucket maximum,
     "text": bf16 (B, 154, 2304), "text_pooled": bf16 (B, 768)}
Reference (receiver, src/model_trainer.py:353-387): three blocking NCCL recvs, then the true latent
shape is read back from the +inf counts of sample 0 / channel 0 (two `.item()` device syncs), the
latents are recovered with a boolean-mask gather (`x[x != inf].reshape(shape)`), and t / the three
null masks are drawn on the CPU.

Here the batch stays in pinned HOST memory until one async H2D copy (mmdit.train.HostFeed):
the shape is inferred on the host from the same +inf counts (no device sync), the padded latents
travel as one contiguous DMA, and the crop is a strided slice + one copy kernel on the device,
bit-identical to the reference's mask gather (tests/test_feed_cpu.py).
"""
import torch
import torch.nn.functional as F

from src.helpers.TimeSampler import TimeSampler

BF16 = torch.bfloat16
INF = float("inf")


def pad_latents(x0, max_side):
    """Sender side (VAE_T5_CLIP.py:438): pad (B,C,h,w) latents to (B,C,max_side,max_side) with +inf."""
    h, w = x0.shape[-2:]
    if h > max_side or w > max_side:
        raise ValueError(f"latent {h}x{w} exceeds the bucket maximum {max_side}")
    return F.pad(x0, (0, max_side - w, 0, max_side - h), value=INF)


def infer_latent_shape(images):
    """(h, w) of the un-padded latents from the +inf counts of sample 0, channel 0
    (model_trainer.py:363-368: rows = H - #inf in column 0, cols = W - #inf in row 0).
    `images` should live on the host (pinned): on a CUDA tensor this is the reference's device sync."""
    plane = images[0, 0]
    isinf = plane == INF
    h = plane.shape[0] - int(isinf[:, 0].sum())
    w = plane.shape[1] - int(isinf[0, :].sum())
    if h <= 0 or w <= 0:
        raise ValueError("wire batch holds no latent data (all +inf)")
    return h, w


def crop_latents(images, h, w):
    """== images[images != inf].reshape(B, C, h, w) (model_trainer.py:370) without the mask gather."""
    return images[:, :, :h, :w].contiguous()


def draw_step_randoms(batch, time_sampler=None, p_null=(0.1, 0.316, 0.316)):
    """t and the three null masks, drawn on the CPU in the reference's order (model_trainer.py:378-387,
    defaults of train.py:53-55)."""
    ts = time_sampler or TimeSampler()
    t = ts(batch)
    pp, pg, pb = torch.rand(batch), torch.rand(batch), torch.rand(batch)
    return dict(t=t, null_pooled=pp < p_null[0], null_gemma=pg < p_null[1], null_bert=pb < p_null[2])


def from_wire(wire, time_sampler=None, p_null=(0.1, 0.316, 0.316), pin=True):
    """Reference wire dict (host tensors) -> (host batch for RFTrainer / HostFeed, (h, w)).
    The latents stay PADDED in the returned batch (`x0_padded`) so that the H2D copy is one contiguous
    transfer; `finish_on_device` crops them after the copy."""
    images, text, pooled = wire["images"], wire["text"], wire["text_pooled"]
    if images.is_cuda:
        raise ValueError("from_wire expects host tensors (the point is to avoid the device round trip)")
    B = images.shape[0]
    if text.shape[0] != B or pooled.shape[0] != B or text.dim() != 3 or pooled.dim() != 2:
        raise ValueError(f"inconsistent wire batch: images {tuple(images.shape)}, text {tuple(text.shape)}, "
                         f"pooled {tuple(pooled.shape)}")
    h, w = infer_latent_shape(images)
    out = dict(x0_padded=images.to(BF16), c=text.to(BF16), pooled=pooled.to(BF16))
    out.update(draw_step_randoms(B, time_sampler, p_null))
    if pin and torch.cuda.is_available():
        out = {k: (v if v.is_pinned() else v.pin_memory()) for k, v in out.items()}
    return out, (h, w)


def finish_on_device(dev_batch, hw):
    """After the H2D copy: replace `x0_padded` by the cropped latents `x0` (device-side slice copy)."""
    b = dict(dev_batch)
    b["x0"] = crop_latents(b.pop("x0_padded"), *hw)
    return b

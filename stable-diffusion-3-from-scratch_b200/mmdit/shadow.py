"""bf16 "shadow" copies of fp32 parameters for the tensor-core kernels.

The reference keeps fp32 parameters and lets torch.autocast cast every weight to
bf16 on every use (model_trainer.py:416); here each module packs the weights a
GEMM needs (e.g. query|key|value) row-wise into one persistent bf16 buffer.

Who refreshes a buffer:
  * FusedAdamW (mmdit/optim.py) rewrites the shadow in the same pass that updates
    the fp32 master weight; such buffers are marked *managed* and never re-cast.
  * otherwise, while gradients are enabled, the buffer is re-cast on every forward
    (the optimizer has just changed the weights; the re-cast is then also recorded
    when a training step is captured into a CUDA graph);
  * under no_grad (sampling) the copy is reused until a parameter's version or
    storage changes.
"""
import torch

BF16 = torch.bfloat16


class Entry:
    __slots__ = ("key", "buf", "params", "managed")

    def __init__(self, key, buf, params):
        self.key, self.buf, self.params, self.managed = key, buf, params, False


def _key(params):
    return tuple((p.data_ptr(), p._version, p.device) for p in params)


def _ptr_key(params):
    return tuple((p.data_ptr(), p.device) for p in params)


def packed_weight(owner, name, params):
    """Row-wise concatenation of 2-D (or conv) weights as one bf16 [sum(n_i), K] tensor."""
    cache = owner.__dict__.setdefault("_mmdit_shadow", {})
    ent = cache.get(name)
    if ent is not None and ent.managed and _ptr_key(ent.params) == _ptr_key(params):
        return ent.buf                      # kept fresh by the fused optimizer
    key = _key(params)
    if ent is not None and ent.key == key and not torch.is_grad_enabled():
        return ent.buf
    rows = sum(p.shape[0] for p in params)
    K = params[0][0].numel()
    if ent is not None and ent.buf.shape == (rows, K) and ent.buf.device == params[0].device:
        buf = ent.buf
    else:
        buf = torch.empty((rows, K), device=params[0].device, dtype=BF16)
    with torch.no_grad():
        r = 0
        for p in params:
            n = p.shape[0]
            buf[r:r + n].copy_(p.reshape(n, K))
            r += n
    ent = Entry(key, buf, list(params))
    cache[name] = ent
    return buf


def shadow_slices(model):
    """{id(param): bf16 view with the parameter's numel} for every packed buffer built so far."""
    out = {}
    for mod in model.modules():
        for ent in mod.__dict__.get("_mmdit_shadow", {}).values():
            r = 0
            for p in ent.params:
                n = p.shape[0]
                out[id(p)] = (ent, ent.buf[r:r + n])
                r += n
    return out

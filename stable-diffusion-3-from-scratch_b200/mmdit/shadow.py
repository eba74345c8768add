"""bf16 "shadow" copies of fp32 parameters for the tensor-core kernels.

The reference keeps fp32 parameters and lets torch.autocast cast every weight to
bf16 on every use (model_trainer.py:416); here each module packs the weights a
GEMM needs (e.g. query|key|value) row-wise into one bf16 buffer.  While
gradients are enabled the buffer is refreshed on every forward (the optimizer
has just changed the weights; the refresh is then also recorded when a training
step is captured into a CUDA graph).  Under no_grad (sampling) the copy is
reused until a parameter's version or storage changes.
"""
import torch

BF16 = torch.bfloat16


def _key(params):
    return tuple((p.data_ptr(), p._version, p.device) for p in params)


def packed_weight(owner, name, params):
    """Row-wise concatenation of 2-D (or conv) weights as one bf16 [sum(n_i), K] tensor."""
    cache = owner.__dict__.setdefault("_mmdit_shadow", {})
    ent = cache.get(name)
    key = _key(params)
    if ent is not None and ent[0] == key and not torch.is_grad_enabled():
        return ent[1]
    rows = sum(p.shape[0] for p in params)
    K = params[0][0].numel()
    buf = ent[1] if ent is not None and ent[1].shape == (rows, K) and ent[1].device == params[0].device \
        else torch.empty((rows, K), device=params[0].device, dtype=BF16)
    with torch.no_grad():
        r = 0
        for p in params:
            n = p.shape[0]
            buf[r:r + n].copy_(p.reshape(n, K))
            r += n
    cache[name] = (key, buf)
    return buf


def packed_bias(owner, name, params):
    """Concatenation of 1-D biases as one fp32 vector (used by the GEMM epilogue)."""
    if len(params) == 1:
        return params[0].detach()
    cache = owner.__dict__.setdefault("_mmdit_shadow", {})
    ent = cache.get(name)
    key = _key(params)
    if ent is not None and ent[0] == key and not torch.is_grad_enabled():
        return ent[1]
    with torch.no_grad():
        buf = torch.cat([p.detach().reshape(-1) for p in params])
    cache[name] = (key, buf)
    return buf

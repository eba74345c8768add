"""bf16 "shadow" copies of fp32 parameters for the tensor-core kernels.

The reference keeps fp32 parameters and lets torch.autocast cast every weight to
bf16 on every use (model_trainer.py:416); here each module packs the weights a
GEMM needs (e.g. query|key|value) row-wise into one persistent bf16 buffer.

Who refreshes a buffer:
  * FusedAdamW (mmdit/optim.py) rewrites the shadow in the same pass that updates
    the fp32 master weight; such buffers are marked *managed*.  The kernel writes the master
    weights through raw pointers, which leaves `Parameter._version` alone -- so a version that
    HAS moved means somebody else wrote the weights in place (load_state_dict, loadModel,
    DeviceEMA.copy_to, a manual re-init) and the shadow is re-cast (into the same storage, so
    descriptor tables and captured graphs that hold its address stay valid).
  * otherwise, while gradients are enabled, the buffer is re-cast on every forward
    (the optimizer has just changed the weights; the re-cast is then also recorded
    when a training step is captured into a CUDA graph);
  * under no_grad (sampling) the copy is reused until a parameter's version or
    storage changes.
"""
import torch

BF16 = torch.bfloat16


class Entry:
    __slots__ = ("key", "buf", "params", "managed", "versions")

    def __init__(self, key, buf, params):
        self.key, self.buf, self.params, self.managed = key, buf, params, False
        self.versions = _versions(params)


def _versions(params):
    return tuple(p._version for p in params)


def _fill(buf, params):
    K = buf.shape[1]
    with torch.no_grad():
        r = 0
        for p in params:
            n = p.shape[0]
            buf[r:r + n].copy_(p.reshape(n, K))
            r += n


class _Cache(dict):
    """Per-module shadow cache.  Derived state: a deep copy of the module (the reference keeps its EMA
    as `copy.deepcopy(model).cpu()`, model_trainer.py:256) or a pickle starts with an empty cache
    instead of dragging device buffers along."""

    def __deepcopy__(self, memo):
        return _Cache()

    def __reduce__(self):
        return (_Cache, ())


def _key(params):
    return tuple((p.data_ptr(), p._version, p.device) for p in params)


def _ptr_key(params):
    return tuple((p.data_ptr(), p.device) for p in params)


def packed_weight(owner, name, params):
    """Row-wise concatenation of 2-D (or conv) weights as one bf16 [sum(n_i), K] tensor."""
    cache = owner.__dict__.setdefault("_mmdit_shadow", _Cache())
    ent = cache.get(name)
    if ent is not None and ent.managed and _ptr_key(ent.params) == _ptr_key(params):
        if ent.versions != _versions(params):     # written in place by someone other than the optimizer
            _fill(ent.buf, params)
            ent.versions = _versions(params)
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
    _fill(buf, params)
    ent = Entry(key, buf, list(params))
    cache[name] = ent
    return buf


def refresh_stale(model):
    """Re-cast every managed shadow whose master weights were written in place since the last
    cast (see the module docstring).  A captured training step replays without running Python, so
    the trainer calls this before every replay (a host-side version compare; no device work unless
    something changed).  Returns the number of buffers re-cast."""
    n = 0
    for mod in model.modules():
        for ent in mod.__dict__.get("_mmdit_shadow", {}).values():
            if ent.managed and ent.versions != _versions(ent.params):
                _fill(ent.buf, ent.params)
                ent.versions = _versions(ent.params)
                n += 1
    return n


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

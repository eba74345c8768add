"""One zero-fill per training step for the small fp32 accumulators of the backward.

The backward of one cfg2 step asks for ~94 zeroed fp32 vectors (bias-gradient column sums, the QK-norm
weight gradients, the SwiGLU bias sums, ...), each a `torch.zeros` = one 2.6 us fill kernel in the
captured step.  `RFTrainer._fwd_bwd` brackets the step with `begin()` / `end()`; in between `zeros()`
hands out 256-byte aligned views of one pre-zeroed buffer, and `begin()` re-zeroes the used extent with a
single fill.  Outside a trainer step (samplers, the reference's own training loop driving the drop-in
modules, tests) nothing changes: `zeros()` is `torch.zeros`.

Invariant: everything above `high` (the largest extent ever handed out) is still zero from the
allocation, so a step that needs more than the previous one stays correct.  Inside a CUDA-graph capture a
view beyond the extent zeroed by the captured `begin()` is refused (a replay would not re-zero it) and the
caller gets `torch.zeros`.  The buffer is never reallocated, so the pointers captured graphs hold stay
valid; requests that do not fit fall back to `torch.zeros`.  The views become `.grad` of parameters: they
stay intact until the next `begin()`, i.e. past the optimizer step of their own training step (the
trainer has no gradient accumulation over several backward passes).  MMDIT_ZERO_POOL=0 disables it."""
import math
import os

import torch

F32 = torch.float32
ALIGN = 64            # floats: every view starts on a 256-byte boundary
POOL_FLOATS = 4 << 20  # 16 MB per device

ENABLED = os.environ.get("MMDIT_ZERO_POOL", "1") != "0"


class _Pool:
    def __init__(self, device):
        self.buf = torch.zeros(POOL_FLOATS, device=device, dtype=F32)
        self.off = self.high = self.zeroed = 0
        self.fills = self.served = self.refused = 0


_pools = {}
_active = None   # pool of the step in progress


def begin(device):
    """Start of a step on `device`: one fill over what earlier steps dirtied."""
    global _active
    if not ENABLED:
        return
    device = torch.device(device)
    key = (device.type, device.index)
    pool = _pools.get(key)
    if pool is None:
        pool = _pools[key] = _Pool(device)
    if pool.high:
        pool.buf[:pool.high].zero_()
        pool.fills += 1
    pool.zeroed = pool.high
    pool.off = 0
    _active = pool


def end():
    global _active
    _active = None


def zeros(shape, device):
    """fp32 zeros of `shape` on `device`: a view of the step's pool when one is active, else torch.zeros."""
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    pool = _active
    if pool is not None and pool.buf.device == torch.device(device):
        n = math.prod(shape)
        stop = pool.off + (n + ALIGN - 1) // ALIGN * ALIGN
        ok = stop <= POOL_FLOATS
        if ok and stop > pool.zeroed and pool.buf.is_cuda and torch.cuda.is_current_stream_capturing():
            ok = False
        if ok:
            view = pool.buf[pool.off:pool.off + n].view(shape)
            pool.off = stop
            pool.high = max(pool.high, stop)
            pool.served += 1
            return view
        pool.refused += 1
    return torch.zeros(shape, device=device, dtype=F32)


def stats(device):
    device = torch.device(device)
    p = _pools.get((device.type, device.index))
    return None if p is None else {"high": p.high, "fills": p.fills, "served": p.served, "refused": p.refused}

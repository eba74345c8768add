"""CPU: the per-step zero pool (mmdit/zeropool.py) hands out zeroed, disjoint, aligned views, re-zeroes
what a step dirtied with one fill, never reallocates, and is plain torch.zeros outside a step."""
import torch

from mmdit import zeropool


def test_outside_a_step_it_is_torch_zeros():
    zeropool.end()
    a = zeropool.zeros((3, 5), "cpu")
    assert a.dtype == torch.float32 and a.shape == (3, 5) and float(a.abs().sum()) == 0
    st = zeropool.stats("cpu")
    assert st is None or a.untyped_storage().data_ptr() != zeropool._pools[("cpu", None)].buf.untyped_storage().data_ptr()


def test_views_are_disjoint_aligned_and_rezeroed_by_one_fill():
    zeropool.begin("cpu")
    pool = zeropool._active
    base, fills0 = pool.buf.data_ptr(), pool.fills
    a = zeropool.zeros(100, "cpu")
    b = zeropool.zeros((4, 64), "cpu")
    assert (a.data_ptr() - base) % 256 == 0 and (b.data_ptr() - base) % 256 == 0
    assert b.data_ptr() >= a.data_ptr() + 100 * 4
    a.fill_(3.0)
    b.fill_(5.0)
    zeropool.end()
    assert float(zeropool.zeros(7, "cpu").sum()) == 0      # not from the pool any more
    high = pool.high
    zeropool.begin("cpu")                                   # next step: one fill over the dirty extent
    assert pool.fills == fills0 + 1 and pool.buf.data_ptr() == base
    a2 = zeropool.zeros(100, "cpu")
    c2 = zeropool.zeros(300, "cpu")                         # a different request sequence is fine
    d2 = zeropool.zeros(5000, "cpu")                        # beyond the old extent: still pristine zeros
    assert a2.data_ptr() == a.data_ptr() and float(a2.sum()) == 0 and float(c2.sum()) == 0 and float(d2.sum()) == 0
    assert pool.high > high
    zeropool.end()


def test_requests_that_do_not_fit_fall_back():
    zeropool.begin("cpu")
    pool = zeropool._active
    big = zeropool.zeros(zeropool.POOL_FLOATS + 1, "cpu")
    assert big.data_ptr() < pool.buf.data_ptr() or big.data_ptr() >= pool.buf.data_ptr() + 4 * zeropool.POOL_FLOATS
    assert pool.refused >= 1 and float(big.sum()) == 0
    zeropool.end()

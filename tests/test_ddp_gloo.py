"""CPU, world_size 2 over gloo: the data-parallel host logic (mmdit/train.py GradBuckets) --
bucket layout by block, hook-driven launch order, and mean all-reduce equal to what DDP
(model_trainer.py:224) computes: 2 ranks x B must reproduce the 1 rank x 2B gradient."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Toy(nn.Module):
    """Same naming scheme as diff_model (blocks.N.* + top-level params) with plain torch math."""

    def __init__(self):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Linear(8, 8) for _ in range(3)])
        self.head = nn.Linear(8, 4)
        self.frozen = nn.Parameter(torch.ones(3), requires_grad=False)
        self.unused = nn.Parameter(torch.ones(5))      # never reached by backward: slot must read 0

    def forward(self, x):
        for b in self.blocks:
            x = torch.tanh(b(x))
        return self.head(x)


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
    from mmdit.train import GradBuckets
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = Toy()
    buckets = GradBuckets(list(model.named_parameters()), world, None, torch.device("cpu"))
    buckets.install_hooks()
    order = []
    orig = buckets._launch
    buckets._launch = lambda bi: (order.append(buckets.buckets[bi][0]), orig(bi))[1]
    g = torch.Generator().manual_seed(123)
    x = torch.randn(2 * world, 8, generator=g)
    y = torch.randn(2 * world, 4, generator=g)
    sl = slice(2 * rank, 2 * rank + 2)
    for _ in range(2):                      # two steps: buckets must be reusable
        buckets.reset()       # .grad dropped; the hooks move each fresh gradient into its bucket slot
        ((model(x[sl]) - y[sl]) ** 2).mean().backward()
        buckets.finish()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}
    assert float(model.unused.grad.abs().max()) == 0.0
    # single-process reference on the full batch
    ref = Toy()
    ref.load_state_dict(model.state_dict())
    ((ref(x) - y) ** 2).mean().backward()
    err = max(float((grads[k] - p.grad).abs().max()) for k, p in ref.named_parameters()
              if p.requires_grad and p.grad is not None)
    views_ok = all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4
                   for _, flat, ps in buckets.buckets for p in ps)
    q.put((rank, err, order, [k for k, _, _ in buckets.buckets], views_ok))
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    for rank, err, order, keys, views_ok in res:
        assert err < 1e-6, (rank, err)
        assert views_ok
        # one bucket per block, last block first (the order backward produces them), rest last
        assert keys == [("block", 2), ("block", 1), ("block", 0), ("rest", 0)]
        assert order[-4:] == [("rest", 0), ("block", 2), ("block", 1), ("block", 0)] or \
            order[-4:] == [("block", 2), ("block", 1), ("block", 0), ("rest", 0)] or len(order) >= 4


def test_bucket_slots_are_aligned_and_packed_groups_are_adjacent():
    """Host logic of the direct-to-bucket weight gradients (no process group needed): every slot
    starts on a 64-byte boundary, the parameters of one packed GEMM (`adjacent` hint) sit back to
    back in the hinted order, and functional._wgrad_slot only aliases slots that really are."""
    sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
    from mmdit.functional import _wgrad_slot
    from mmdit.train import GradBuckets

    class Blk(nn.Module):
        def __init__(self):
            super().__init__()
            self.s = nn.Parameter(torch.zeros(3))            # odd size: forces padding after it
            self.q = nn.Linear(16, 16, bias=False)
            self.b = nn.Parameter(torch.zeros(5))
            self.k = nn.Linear(16, 16, bias=False)
            self.v = nn.Linear(16, 16, bias=False)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = nn.ModuleList([Blk(), Blk()])
            self.head = nn.Linear(16, 4)

    m = M()
    hint = [[blk.v.weight, blk.q.weight, blk.k.weight] for blk in m.blocks]   # packed order != named order
    gb = GradBuckets(list(m.named_parameters()), 1, None, torch.device("cpu"), peer=False, adjacent=hint)
    assert [k for k, _, _ in gb.buckets] == [("block", 1), ("block", 0), ("rest", 0)]
    for _, flat, ps in gb.buckets:
        assert {id(p) for p in ps} <= {id(p) for p in m.parameters()}
        for p in ps:
            assert p._grad_slot.data_ptr() % 64 == 0 and p._grad_slot.shape == p.shape
            assert flat.data_ptr() <= p._grad_slot.data_ptr() < flat.data_ptr() + flat.numel() * 4
    for blk in m.blocks:
        out = _wgrad_slot([blk.v.weight, blk.q.weight, blk.k.weight])
        assert out is not None and out.shape == (48, 16)
        out.fill_(0.0)
        out[16:32].fill_(2.0)                                   # rows of the 2nd packed parameter
        assert float(blk.q.weight._grad_slot.sum()) == 2.0 * 256 and float(blk.v.weight._grad_slot.sum()) == 0.0
        assert _wgrad_slot([blk.q.weight, blk.v.weight]) is None   # not adjacent in this order
    assert _wgrad_slot([nn.Parameter(torch.zeros(4, 4))]) is None  # no slot at all
    # a step without any backward: every slot reads zero, .grad is the slot
    gb.reset()
    gb.all_reduce_mean()
    assert all(p.grad is p._grad_slot and float(p.grad.abs().sum()) == 0.0 for p in m.parameters())

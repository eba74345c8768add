"""CPU: bf16 shadow-weight bookkeeping (mmdit/shadow.py).  A shadow that the fused optimizer manages
must still notice in-place writes by anybody else -- load_state_dict / loadModel (diff_model.py:553-579),
DeviceEMA.copy_to, manual re-initialisation -- and must keep its storage (descriptor tables and
captured graphs hold its address)."""
import torch
from torch import nn

from mmdit import shadow


def _managed(lin):
    buf = shadow.packed_weight(lin, "w", [lin.weight])
    lin.__dict__["_mmdit_shadow"]["w"].managed = True      # what FusedAdamW.bind_shadows does
    return buf


def test_managed_shadow_follows_load_state_dict_and_keeps_its_storage():
    torch.manual_seed(0)
    lin, other = nn.Linear(16, 8), nn.Linear(16, 8)
    buf = _managed(lin)
    assert torch.equal(buf, lin.weight.detach().bfloat16())
    lin.load_state_dict(other.state_dict())                # in-place copy_: bumps Parameter._version
    with torch.enable_grad():
        again = shadow.packed_weight(lin, "w", [lin.weight])
    assert again.data_ptr() == buf.data_ptr()
    assert torch.equal(again, other.weight.detach().bfloat16())


def test_managed_shadow_is_left_alone_when_only_raw_pointer_writes_happened():
    """The fused optimizer writes master weight AND shadow through raw pointers (no version bump):
    packed_weight must then hand the buffer back untouched (no re-cast pass per forward)."""
    lin = nn.Linear(16, 8)
    buf = _managed(lin)
    buf.fill_(3.0)                                         # stands for the kernel's own refresh
    assert float(shadow.packed_weight(lin, "w", [lin.weight]).float().mean()) == 3.0


def test_refresh_stale_recasts_only_what_changed():
    net = nn.Sequential(nn.Linear(8, 8), nn.Linear(8, 4))
    bufs = [_managed(m) for m in net]
    assert shadow.refresh_stale(net) == 0
    with torch.no_grad():
        net[1].weight.mul_(2.0)
    assert shadow.refresh_stale(net) == 1
    assert torch.equal(bufs[1], net[1].weight.detach().bfloat16())
    assert shadow.refresh_stale(net) == 0

"""CPU: DeviceEMA == the reference's EMA loop (model_trainer.py:537-541), bit for bit, and its
state_dict is what saveModel(EMA_state_dict=...) expects (the model's keys and shapes)."""
import copy

import torch
from torch import nn

from mmdit.ema import DeviceEMA


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(8, 8)
        self.b = nn.Linear(8, 3, bias=False)
        self.frozen = nn.Parameter(torch.arange(4.0), requires_grad=False)
        self.register_buffer("buf", torch.ones(2))


def test_device_ema_matches_reference_loop_bitwise():
    torch.manual_seed(0)
    model = Toy()
    ref = copy.deepcopy(model)                       # model_trainer.py:256
    ema = DeviceEMA(model, decay=0.99, update_freq=3)
    for step in range(1, 10):
        with torch.no_grad():
            for p in model.parameters():
                if p.requires_grad:
                    p.add_(torch.randn_like(p) * 0.1)   # "an optimizer step"
        did = ema.update(step)
        assert did == (step % 3 == 0)
        if step % 3 == 0:                            # model_trainer.py:537-541, verbatim arithmetic
            with torch.no_grad():
                for ema_param, param in zip(ref.parameters(), model.parameters()):
                    if param.requires_grad:
                        ema_param.data.mul_(0.99).add_(param.data, alpha=(1.0 - 0.99))
    sd, rsd = ema.state_dict(), ref.state_dict()
    assert list(sd) == list(rsd) == list(model.state_dict())
    for k in sd:
        assert sd[k].shape == rsd[k].shape and torch.equal(sd[k], rsd[k]), k
    fresh = Toy()
    ema.copy_to(fresh)
    assert all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), sd.values()))

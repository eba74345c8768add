"""Device-resident EMA of the model weights.

Reference: the EMA lives in a deep copy of the model on the CPU (model_trainer.py:256) and every
`ema_update_freq` steps every parameter is copied D2H and blended there (:537-541) -- a
multi-second stall at 2.5 B parameters.  Here the EMA weights stay on the model's device (one more
fp32 copy: 10 GB of 180 GB at cfg3) and an update is two multi-tensor passes with the reference's
arithmetic in the reference's order (`ema.mul_(decay).add_(p, alpha=1-decay)`), so the result is
bit-identical to the reference formula on the same device.  `state_dict()` has the keys and
shapes `diff_model.saveModel(EMA_state_dict=...)` writes to `model_ema_*.pkl` (:525-526).
"""
import torch


class DeviceEMA:
    def __init__(self, model, decay=0.99, update_freq=100):
        """decay / update_freq: train.py:58-59 defaults."""
        self.decay, self.update_freq = float(decay), int(update_freq)
        self.model = model
        self.names = [n for n, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in model.named_parameters() if p.requires_grad]
        with torch.no_grad():
            self.ema = [p.detach().clone() for p in self.params]

    @torch.no_grad()
    def update(self, step=None):
        """Blend the current weights in.  With `step` given, only every `update_freq`-th step does
        (model_trainer.py:537); returns whether an update happened."""
        if step is not None and step % self.update_freq != 0:
            return False
        torch._foreach_mul_(self.ema, self.decay)
        torch._foreach_add_(self.ema, [p.detach() for p in self.params], alpha=1.0 - self.decay)
        return True

    def state_dict(self):
        """The model's state_dict with the trainable entries replaced by their EMA (frozen entries,
        e.g. the rotary frequencies, are carried over as the reference's deep copy does)."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        for n, e in zip(self.names, self.ema):
            sd[n] = e.detach().clone()
        return sd

    def load_state_dict(self, sd):
        with torch.no_grad():
            for n, e in zip(self.names, self.ema):
                e.copy_(sd[n])

    def copy_to(self, model):
        """Load the EMA weights into `model` (for sampling with the averaged weights)."""
        model.load_state_dict(self.state_dict(), strict=True)

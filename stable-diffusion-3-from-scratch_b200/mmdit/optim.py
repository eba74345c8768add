"""Fused optimizer step for the MMDiT trainer: global-norm clip + AdamW + bf16 shadow refresh
in two kernels over all parameters (reference: model_trainer.py:260 optimizer settings,
:483-503 unscale / clip_grad_norm_(1.0) / step / zero_grad).  State lives in fp32 torch tensors
(exp_avg, exp_avg_sq) so checkpoints keep torch.optim.AdamW's state_dict layout."""
import ctypes as C

import torch

from . import _lib
from .shadow import shadow_slices

F32 = torch.float32


class ParamDesc(C.Structure):
    """Mirror of mmdit_param_desc."""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("shadow", C.c_void_p), ("n", C.c_int64)]


class FusedAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=1.0):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        dev = self.params[0].device
        self.device = dev
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.state = torch.zeros(2, device=dev, dtype=F32)      # [grad sum of squares, step]
        chunk = _lib.lib().mmdit_adamw_chunk_elems()
        work = [(ti, ci) for ti, p in enumerate(self.params) for ci in range((p.numel() + chunk - 1) // chunk)]
        self.n_chunks = len(work)
        self.chunks = torch.tensor(work, dtype=torch.int32).to(dev)
        n = len(self.params)
        self.host_table = torch.zeros(n * C.sizeof(ParamDesc), dtype=torch.uint8).pin_memory()
        self.dev_table = torch.zeros(n * C.sizeof(ParamDesc), dtype=torch.uint8, device=dev)
        self._descs = (ParamDesc * n).from_address(self.host_table.data_ptr())
        self._last = None
        self._shadow = {}

    def bind_shadows(self):
        """Take over every bf16 shadow buffer the modules have built so far (call after a forward)."""
        self._shadow = shadow_slices(self.model)
        for ent, _ in self._shadow.values():
            ent.managed = True
        self._last = None

    @torch.no_grad()
    def step(self):
        key = tuple(p.grad.data_ptr() for p in self.params)
        if key != self._last:      # gradient storage moved (eager mode): refresh the descriptor table
            for i, p in enumerate(self.params):
                g = p.grad
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                d = self._descs[i]
                d.p, d.g, d.m, d.v = p.data_ptr(), g.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr()
                sh = self._shadow.get(id(p))
                d.shadow = sh[1].data_ptr() if sh is not None else None
                d.n = p.numel()
            self.dev_table.copy_(self.host_table, non_blocking=True)
            self._last = tuple(p.grad.data_ptr() for p in self.params)
        _lib.check(_lib.lib().mmdit_adamw_step(
            self.dev_table.data_ptr(), self.chunks.data_ptr(), self.n_chunks, self.state.data_ptr(),
            self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.max_norm,
            torch.cuda.current_stream().cuda_stream), "mmdit_adamw_step")

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def grad_norm(self):
        """Global gradient norm seen by the last step (before clipping)."""
        return self.state[0].sqrt()

    def state_dict(self):
        return {"state": {i: {"step": self.state[1].clone(), "exp_avg": m, "exp_avg_sq": v}
                          for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))},
                "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps,
                                  "weight_decay": self.wd, "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        for i, st in sd["state"].items():
            self.exp_avg[int(i)].copy_(st["exp_avg"])
            self.exp_avg_sq[int(i)].copy_(st["exp_avg_sq"])
            self.state[1] = float(st["step"])

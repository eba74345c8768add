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


class FusedAdamW(torch.optim.Optimizer):
    """Drop-in for the reference's `torch.optim.AdamW(params, lr, eps=1e-8, weight_decay=0.01)`
    (model_trainer.py:260) plus the `clip_grad_norm_(1.0)` in front of it (:487).

    A real `torch.optim.Optimizer`: `param_groups`, `state`, `state_dict()` / `load_state_dict()`
    have torch.optim.AdamW's layout (optim_*.pkl written by saveModel, diff_model.py:527-528, loads
    into either class), and `torch.optim.lr_scheduler.*` / transformers' `get_*_schedule_with_warmup`
    (model_trainer.py:25-41) drive it unchanged: they write `param_groups[0]["lr"]`, and `step()`
    forwards that value to the device-resident learning rate (state[2]) -- a by-value argument would
    be frozen into a captured CUDA graph.  Call `sync_lr()` after `scheduler.step()` when replaying
    a captured step (RFTrainer does)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=1.0):
        self.model = model
        # like the reference (`AdamW(self.model.parameters())`, model_trainer.py:260) the group lists EVERY
        # parameter, frozen ones included (they never get state), so state_dict indices line up
        every = list(model.parameters())
        params = [p for p in every if p.requires_grad]
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                        maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=True)
        super().__init__(every, defaults)
        pos = {id(p): j for j, p in enumerate(params)}
        self._index = {i: pos[id(p)] for i, p in enumerate(every) if p.requires_grad}   # group index -> trainable index
        self.params = params
        self.max_norm = max_norm
        dev = self.params[0].device
        self.device = dev
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.dstate = torch.zeros(4, device=dev, dtype=F32)     # [grad sum of squares, step, lr, -]
        for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq):
            # torch.optim.AdamW's per-parameter state; `step` is one shared device scalar
            self.state[p] = {"step": self.dstate[1], "exp_avg": m, "exp_avg_sq": v}
        self._lr_dev = None
        self._lr_host = torch.zeros(1, dtype=F32)
        if dev.type == "cuda":
            self._lr_host = self._lr_host.pin_memory()
        chunk = _lib.lib().mmdit_adamw_chunk_elems() if dev.type == "cuda" else 16384
        work = [(ti, ci) for ti, p in enumerate(self.params) for ci in range((p.numel() + chunk - 1) // chunk)]
        self.n_chunks = len(work)
        self.chunks = torch.tensor(work, dtype=torch.int32).to(dev)
        n = len(self.params)
        self.host_table = torch.zeros(n * C.sizeof(ParamDesc), dtype=torch.uint8)
        if dev.type == "cuda":
            self.host_table = self.host_table.pin_memory()
        self.dev_table = torch.zeros(n * C.sizeof(ParamDesc), dtype=torch.uint8, device=dev)
        self._descs = (ParamDesc * n).from_address(self.host_table.data_ptr())
        self._last = None
        self._shadow = {}
        self.sync_lr()

    # attribute-style access to the hyper-parameters (they live in param_groups[0])
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = float(v)

    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    wd = property(lambda self: self.param_groups[0]["weight_decay"])

    def sync_lr(self):
        """Make the device-resident learning rate equal param_groups[0]["lr"] (tiny async H2D copy
        on the current stream; skipped when unchanged)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_dev and not (self.device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            self._lr_host[0] = lr
            self.dstate[2:3].copy_(self._lr_host, non_blocking=True)
            self._lr_dev = lr

    def warm_kernels(self):
        """Launch the optimizer kernels once on a throw-away 16-element parameter: loads the module
        ahead of a CUDA-graph capture without touching the model or this optimizer's state."""
        if self.device.type != "cuda":
            return
        t = torch.zeros(5, 16, device=self.device, dtype=F32)          # p, g, m, v, state
        table = torch.zeros(C.sizeof(ParamDesc), dtype=torch.uint8)
        d = ParamDesc.from_address(table.data_ptr())
        d.p, d.g, d.m, d.v, d.shadow, d.n = (t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                                             t[3].data_ptr(), None, 16)
        dt = table.to(self.device)
        chunks = torch.zeros(1, 2, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().mmdit_adamw_step(dt.data_ptr(), chunks.data_ptr(), 1, t[4].data_ptr(), 0.0, 0.9,
                                               0.999, 1e-8, 0.0, 1.0, torch.cuda.current_stream().cuda_stream),
                   "mmdit_adamw_step")
        torch.cuda.current_stream().synchronize()     # `t`, `dt`, `chunks` die with this frame

    def bind_shadows(self):
        """Take over every bf16 shadow buffer the modules have built so far (call after a forward)."""
        self._shadow = shadow_slices(self.model)
        for ent, _ in self._shadow.values():
            ent.managed = True
        self._last = None

    @torch.no_grad()
    def step(self, closure=None):
        self.sync_lr()
        key = tuple(0 if p.grad is None else p.grad.data_ptr() for p in self.params)
        if key != self._last:      # gradient storage moved (eager mode): refresh the descriptor table
            for i, p in enumerate(self.params):
                g = p.grad
                if g is None:      # no gradient reached this parameter: torch.optim skips it; a zero
                    g = p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)   # gradient only decays it
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                d = self._descs[i]
                d.p, d.g, d.m, d.v = p.data_ptr(), g.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr()
                sh = self._shadow.get(id(p))
                d.shadow = sh[1].data_ptr() if sh is not None else None
                d.n = p.numel()
            self.dev_table.copy_(self.host_table, non_blocking=True)
            self._last = tuple(p.grad.data_ptr() for p in self.params)
        g0 = self.param_groups[0]
        _lib.check(_lib.lib().mmdit_adamw_step(
            self.dev_table.data_ptr(), self.chunks.data_ptr(), self.n_chunks, self.dstate.data_ptr(),
            -1.0, g0["betas"][0], g0["betas"][1], g0["eps"], g0["weight_decay"], self.max_norm,
            torch.cuda.current_stream().cuda_stream), "mmdit_adamw_step")

    def grad_norm(self):
        """Global gradient norm seen by the last step (before clipping)."""
        return self.dstate[0].sqrt()

    def step_count(self):
        return self.dstate[1]

    def load_state_dict(self, sd):
        """Accepts a torch.optim.AdamW (or FusedAdamW) state_dict; the moments are copied into this
        optimizer's own buffers (the kernel's descriptor table and captured graphs hold their addresses)."""
        for i, st in sd["state"].items():
            j = self._index[int(i)]
            self.exp_avg[j].copy_(st["exp_avg"])
            self.exp_avg_sq[j].copy_(st["exp_avg_sq"])
            self.dstate[1] = float(st["step"])
        if sd.get("param_groups"):
            g = sd["param_groups"][0]
            for k in ("lr", "betas", "eps", "weight_decay", "initial_lr"):
                if k in g:
                    self.param_groups[0][k] = tuple(g[k]) if k == "betas" else g[k]
        self.sync_lr()

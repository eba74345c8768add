"""Rectified-flow training step around diff_model: the body of the reference's hot loop
(src/model_trainer.py:378-503) without the loader-rank receive, wandb and EMA plumbing.

    t ~ logit-normal (TimeSampler), null masks, x_t = (1-t) x0 + t eps, v = model(...),
    loss = mean((v - (eps - x0))^2), backward, [gradient all-reduce], clip 1.0, AdamW, zero_grad

One process per GPU.  Data parallelism (model_trainer.py:224) is a bucketed fp32 gradient
all-reduce (mean); buckets follow the block structure and are launched on a side stream as soon
as a block's gradients are final, so the reduction of block i overlaps the backward of block
i-1.  On CUDA the exchange is our own peer-memory kernel (mmdit/comm.py, csrc/comm.cu: one
launch per bucket, reduce-scatter + all-gather + mean over NVLink), which is captured together
with everything else: with `use_graph=True` the whole step (noise -> forward -> backward with
the overlapped bucket reductions -> clip + AdamW) is ONE CUDA graph at any world size.
torch.distributed collectives (NCCL / gloo) remain as the fallback exchange (`peer=False`).
"""
import torch
import torch.distributed as dist

from . import ops, shadow, streams, zeropool
from .functional import rf_loss
from .optim import FusedAdamW

BF16, F32 = torch.bfloat16, torch.float32


def host_batch(B, C, h, w, M=154, class_dim=768, seed=0, p_null=(0.1, 0.316, 0.316), pin=True):
    """Synthetic batch in the reference's wire format (model_trainer.py:353-355): bf16 latents,
    bf16 text (B,154,2304), bf16 pooled (B,768); t and the null masks are drawn on the CPU as
    the trainer does (:378-387).  Tensors are pinned for async H2D copies."""
    g = torch.Generator().manual_seed(seed)
    out = dict(
        x0=torch.randn((B, C, h, w), generator=g).to(BF16),
        c=torch.randn((B, M, 2304), generator=g).to(BF16),
        pooled=torch.randn((B, class_dim), generator=g).to(BF16),
        t=torch.sigmoid(torch.randn(B, generator=g)),
        null_pooled=torch.rand(B, generator=g) < p_null[0],
        null_gemma=torch.rand(B, generator=g) < p_null[1],
        null_bert=torch.rand(B, generator=g) < p_null[2],
    )
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


class GradBuckets:
    """Flat fp32 gradient buckets for the data-parallel exchange (replaces the DDP reducer,
    model_trainer.py:224).  One bucket per transformer block plus one for everything else.  Every
    parameter owns a 64-byte aligned slot (`p._grad_slot`) inside its bucket; at the start of a
    step `.grad` is dropped, the packed wgrad GEMMs write their slots directly, any other fresh
    gradient is copied into its slot by the post-accumulate hook, and `.grad` then IS the slot --
    no zero-fill and no accumulate pass over the buckets (torch DDP's default copies every gradient
    into its buckets: an extra 4*P-byte pass).  On CUDA the buckets live in one IPC-exported arena
    (mmdit/comm.py) and are reduced by our peer-memory kernel; otherwise by dist.all_reduce."""

    def __init__(self, named_params, world_size, process_group=None, device=None, peer=None, adjacent=()):
        """adjacent: lists of parameters that one packed GEMM produces the gradients of (e.g. the
        adaLN projections of a block); they are laid out back to back, in that order, so that the
        wgrad GEMM can write the bucket directly (functional._wgrad_slot)."""
        self.world_size = world_size
        self.group = process_group
        self.arena = None
        self.overlap = False
        self.main_stream = None
        groups = {}
        for name, p in named_params:
            if not p.requires_grad:
                continue
            groups.setdefault(self.bucket_key(name), []).append(p)
        # backward reaches the LAST block first: reduce in that order
        self.order = sorted(groups, key=lambda k: (k[0] != "block", -k[1]))
        first = {id(g[0]): g for g in adjacent if g}
        member = {id(p) for g in adjacent for p in g}
        for key, ps in groups.items():
            laid = []
            for p in ps:
                if id(p) in first:
                    laid.extend(first[id(p)])
                elif id(p) not in member:
                    laid.append(p)
            if len(laid) == len(ps) and {id(p) for p in laid} == {id(p) for p in ps}:
                groups[key] = laid
        self.buckets = []
        self.ranges = []
        dev = device or groups[self.order[0]][0].device
        # every parameter's slot starts on a 64-byte boundary (vector loads/stores, TMA-free kernels)
        sizes = [sum((p.numel() + 15) // 16 * 16 for p in groups[key]) for key in self.order]
        padded = [(n + 63) // 64 * 64 for n in sizes]      # buckets start 256-byte aligned
        if peer is None:
            peer = (world_size in (2, 4, 8) and torch.device(dev).type == "cuda"
                    and dist.is_available() and dist.is_initialized())
        if peer:
            from .comm import PeerArena
            self.arena = PeerArena(sum(padded), dev, process_group)
            self.side = torch.cuda.Stream(device=dev)
        start = 0
        for key, n, npad in zip(self.order, sizes, padded):
            ps = groups[key]
            if self.arena is not None:
                flat = self.arena.flat[start:start + n]
            else:
                flat = torch.zeros(n, device=dev, dtype=F32)
            self.ranges.append((start, npad))
            start += npad
            off = 0
            for p in ps:
                p._grad_slot = flat[off:off + p.numel()].view_as(p)
                p.grad = None
                off += (p.numel() + 15) // 16 * 16
            self.buckets.append((key, flat, ps))
        self._seen = set()

    @staticmethod
    def bucket_key(name):
        """Bucket of a parameter: its transformer block, or "rest".  The per-block y projections
        (blocks.i.y_proj.*) belong to "rest": diff_model.forward computes them for ALL blocks in one
        GEMM ahead of block 0, so their gradient lands at the very end of the backward -- inside a
        block bucket they would hold every bucket's exchange back until then."""
        parts = name.split(".")
        if parts[0] == "blocks" and parts[2] != "y_proj":
            return ("block", int(parts[1]))
        return ("rest", 0)

    def zero(self):
        if self.arena is not None:
            self.arena.flat.zero_()
            return
        for _, flat, _ in self.buckets:
            flat.zero_()

    def reset(self):
        """Start of a step: drop every .grad.  The backward then either writes a parameter's bucket
        slot directly (packed wgrad GEMMs) or autograd hands over a fresh tensor that the hook
        copies into the slot -- the buckets are never zero-filled or accumulated into."""
        for _, _, ps in self.buckets:
            for p in ps:
                p.grad = None
        self.begin_step()

    def _adopt(self, p):
        """.grad of `p` is final: make it live in its bucket slot."""
        slot = p._grad_slot
        if p.grad is None:
            slot.zero_()
        elif p.grad.data_ptr() != slot.data_ptr():
            if streams.async_wgrad and p.dim() >= 2:   # may still be in flight on the wgrad stream
                streams.join_wgrad(p.device)
            slot.copy_(p.grad)
            self.copied_elems += p.numel()
        else:
            self.direct_elems += p.numel()
        p.grad = slot

    # ---- overlap: launch a bucket's all-reduce the moment its last gradient has landed
    def install_hooks(self):
        self.overlap = True
        self._count = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        for bi, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                p.register_post_accumulate_grad_hook(lambda _p, bi=bi: self._ready(bi, _p))

    def begin_step(self):
        self._count = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._seen = set()
        self._hook_streams = [set() for _ in self.buckets]
        self.direct_elems = self.copied_elems = 0     # gradients written in place / copied into their slot
        if self.arena is not None:
            # the stream the backward runs on (inside a capture: the capturing stream); the hooks
            # fire on autograd's thread, so it is pinned here rather than looked up there
            self.main_stream = torch.cuda.current_stream()

    def _launch(self, bi):
        if self.arena is not None:
            # fork: the side stream picks up after everything issued so far (this bucket's
            # gradients), then runs the bucket's exchange kernel next to the rest of the backward
            # (the gradients of a bucket may come from the caller's stream, from the text-branch
            # stream of the two-stream block schedule, and from the stream a hook copied on)
            srcs = {self.main_stream or torch.cuda.current_stream(), torch.cuda.current_stream()}
            srcs |= self._hook_streams[bi]
            if streams.ENABLED:
                srcs.add(streams.side(self.side.device))
            if streams.async_wgrad:
                srcs.add(streams.wgrad(self.side.device))
            for st in srcs:
                ev = torch.cuda.Event()
                ev.record(st)
                self.side.wait_event(ev)
            off, n = self.ranges[bi]
            self.arena.all_reduce_mean(off, n, stream=self.side)
            self._works[bi] = True
            return
        flat = self.buckets[bi][1]
        self._works[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _ready(self, bi, p):
        with torch.no_grad():
            self._adopt(p)
        self._seen.add(id(p))
        if self.arena is not None:
            self._hook_streams[bi].add(torch.cuda.current_stream())
        self._count[bi] += 1
        if self.overlap and self.world_size > 1 and self._count[bi] == len(self.buckets[bi][2]):
            self._launch(bi)

    def _adopt_missing(self):
        """Parameters no gradient reached this step (or every one, when no hooks are installed)."""
        with torch.no_grad():
            for _, _, ps in self.buckets:
                for p in ps:
                    if id(p) not in self._seen:
                        self._adopt(p)
                        self._seen.add(id(p))

    def all_reduce_mean(self):
        """Non-overlapped variant: reduce every bucket now (async launches, then wait + mean)."""
        self._adopt_missing()
        if self.world_size == 1:
            return
        if self.arena is not None:
            for off, n in self.ranges:
                self.arena.all_reduce_mean(off, n)
            return
        works = [dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for _, flat, _ in self.buckets]
        for w, (_, flat, _) in zip(works, self.buckets):
            w.wait()
            flat.mul_(1.0 / self.world_size)

    def finish(self):
        """Wait for every bucket (launching the ones no hook fired for) and apply DDP's mean."""
        self._adopt_missing()
        if self.world_size == 1:
            return
        for bi in range(len(self.buckets)):
            if self._works[bi] is None:
                self._launch(bi)
        if self.arena is not None:
            # join: the optimizer (current stream) runs after the last exchange kernel
            ev = torch.cuda.Event()
            ev.record(self.side)
            torch.cuda.current_stream().wait_event(ev)
            return
        for bi, (_, flat, _) in enumerate(self.buckets):
            self._works[bi].wait()
            flat.mul_(1.0 / self.world_size)


class HostFeed:
    """Double-buffered host->device feed: the pinned batch of step i+1 is copied on a side stream
    while step i computes (the reference blocks on three NCCL recvs per step, model_trainer.py:353-362)."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.pending = None

    def submit(self, host_batch_, latent_hw=None):
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host_batch_.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = (dev, ev, latent_hw)

    def submit_wire(self, wire, time_sampler=None, p_null=(0.1, 0.316, 0.316)):
        """Queue a batch in the reference's loader wire format (mmdit/feed.py): the latent shape is
        inferred on the host, the +inf-padded latents travel as one contiguous copy."""
        from . import feed
        hb, hw = feed.from_wire(wire, time_sampler, p_null)
        self.submit(hb, hw)

    def take(self):
        dev, ev, hw = self.pending
        torch.cuda.current_stream().wait_event(ev)
        for v in dev.values():
            v.record_stream(torch.cuda.current_stream())
        self.pending = None
        if hw is not None:
            from . import feed
            dev = feed.finish_on_device(dev, hw)     # crop the padded latents (device slice copy)
        return dev


class RFTrainer:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, clip=1.0,
                 world_size=1, process_group=None, use_graph=False, fused_optimizer=True, peer=None,
                 ema=None):
        """ema: optional mmdit.ema.DeviceEMA, blended in after the optimizer every `update_freq`
        steps (model_trainer.py:537-541) -- on the device, no D2H copy of the weights."""
        self.model = model
        self.device = next(model.parameters()).device
        self.clip = clip
        self.world_size = world_size
        self.use_graph = use_graph
        self.params = [p for p in model.parameters() if p.requires_grad]
        adjacent = [m._mod_weights() for m in model.modules() if hasattr(m, "_mod_weights")]
        self.buckets = GradBuckets(list(model.named_parameters()), world_size, process_group,
                                   self.device, peer=peer, adjacent=adjacent) if world_size > 1 else None
        if self.buckets is not None:
            self.buckets.install_hooks()
        self.fused_optimizer = fused_optimizer
        if fused_optimizer:
            # clip + AdamW + bf16 shadow refresh in two kernels (mmdit/optim.py)
            self.opt = FusedAdamW(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=clip)
        else:
            self.opt = torch.optim.AdamW(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                         fused=True, capturable=self.use_graph)
        self.ema = ema
        self.steps_done = 0
        self._bound = False
        self._graphs = {}          # batch geometry -> (graph, graph_opt, static inputs, loss)
        self.graph = None
        self.graph_opt = None
        self.static = None
        self.loss = None
        self.kernel_launches = None

    # ------------------------------------------------------------------ data
    def to_device(self, hb):
        """H2D copy of one host batch (pinned -> device, async on the current stream)."""
        return {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}

    def h2d_bytes(self, hb):
        return sum(v.numel() * v.element_size() for v in hb.values())

    # ------------------------------------------------------------------ step
    def _fwd_bwd(self, b):
        zeropool.begin(b["x0"].device)     # one fill for the backward's small fp32 accumulators
        try:
            return self._fwd_bwd_body(b)
        finally:
            zeropool.end()

    def _fwd_bwd_body(self, b):
        eps = torch.randn_like(b["x0"])                               # diff_model.py:235
        x_t = ops.rf_noise(b["x0"], eps, b["t"])                      # :238
        v = self.model(x_t, b["t"], b["c"], b["pooled"], b["null_pooled"], b["null_gemma"], b["null_bert"])
        loss = rf_loss(v, eps, b["x0"])                               # model_trainer.py:429-446
        streams.keepalive.clear()          # operands of the previous step's weight-gradient GEMMs
        streams.async_wgrad = streams.ASYNC_WGRAD and loss.is_cuda
        try:
            loss.backward()
        finally:
            streams.async_wgrad = False
        if loss.is_cuda:
            streams.join_wgrad(loss.device)   # weight gradients are complete for whoever reads them next
        return loss.detach()

    def _update(self):
        if self.buckets is not None:
            self.buckets.finish()
        if self.fused_optimizer:
            if not self._bound:        # the first forward has built the bf16 shadow buffers
                self.opt.bind_shadows()
                self._bound = True
            self.opt.step()                                           # :483-491 in one pass
        else:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)   # :487
            self.opt.step()                                           # :491

    def _zero(self):
        if self.buckets is not None:
            self.buckets.reset()
        else:
            self.opt.zero_grad(set_to_none=True)                      # :503

    def step(self, batch):
        """batch: dict of DEVICE tensors (see to_device).  Returns the loss (device scalar)."""
        if not self.use_graph:
            self._zero()
            loss = self._fwd_bwd(batch)
            self._update()
            self._after_step()
            return loss
        # one captured step per input geometry: aspect-ratio buckets (feed.from_wire hands over a
        # different latent h x w per bucket, dataset_utils.py:119-161) each get their own graph
        self.graph, self.graph_opt, self.static, self.loss = self.capture(batch)
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)
        # a replay runs no Python: bring the device-side copies of host-managed state up to date
        if self.fused_optimizer:
            shadow.refresh_stale(self.model)      # weights written in place since the last cast (load_state_dict, EMA copy_to)
            self.opt.sync_lr()                    # a scheduler's new param_groups[0]["lr"] (model_trainer.py:495-496)
        self.graph.replay()
        if self.graph_opt is not None:
            # data parallel: the all-reduce sits between the two captured halves of the step
            # (capturing NCCL work issued from autograd hooks hung in testing on this stack)
            self.buckets.all_reduce_mean()
            self.graph_opt.replay()
        self._after_step()
        return self.loss

    def capture(self, batch):
        """Build (once) the captured step for this batch geometry without running it; `step` does this
        lazily.  Capturing does not train and leaves the optimizer state alone."""
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.items()))
        cap = self._graphs.get(key)
        if cap is None:
            cap = self._graphs[key] = self._capture(batch)
        return cap

    def _after_step(self):
        self.steps_done += 1
        if self.ema is not None:
            self.ema.update(self.steps_done)

    def _capture(self, batch):
        """Capture one training step for this batch geometry.  The warm-up that CUDA graphs need
        (allocator, lazy module loading) must not train: with the fused optimizer it runs forward +
        backward (+ exchange) only and the optimizer kernels are loaded by a dry launch on a dummy
        parameter; with torch.optim the parameters and optimizer state are snapshotted and restored."""
        static = {k: v.clone() for k, v in batch.items()}
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        snap = None
        if not self.fused_optimizer:
            snap = ([p.detach().clone() for p in self.params], _clone_state(self.opt.state_dict()))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._zero()
                self._fwd_bwd(static)
                if self.fused_optimizer:
                    if self.buckets is not None:
                        self.buckets.finish()
                else:
                    self._update()
            if self.fused_optimizer:
                self.opt.warm_kernels()
                if not self._bound:
                    self.opt.bind_shadows()
                    self._bound = True
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if snap is not None:
            with torch.no_grad():
                for p, q in zip(self.params, snap[0]):
                    p.copy_(q)
            self.opt.load_state_dict(snap[1])
        self._zero()
        graph = torch.cuda.CUDAGraph()
        graph_opt = None
        if self.buckets is None:
            with torch.cuda.graph(graph):
                loss = self._fwd_bwd(static)
                self._update()
            return graph, None, static, loss
        if self.buckets.arena is not None:
            # peer-memory exchange: plain kernels on a forked side stream, so the overlapped
            # bucket reductions are captured with the rest of the step into one graph
            with torch.cuda.graph(graph):
                self.buckets.reset()
                loss = self._fwd_bwd(static)
                self._update()
            return graph, None, static, loss
        self.buckets.overlap = False            # hooks stay silent while capturing / replaying
        with torch.cuda.graph(graph):
            self.buckets.reset()
            loss = self._fwd_bwd(static)
            self.buckets._adopt_missing()
        self.buckets.all_reduce_mean()
        graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_opt):
            self.opt.step() if self.fused_optimizer else self._update()
        return graph, graph_opt, static, loss


def _clone_state(sd):
    """Deep copy of a torch optimizer state_dict (tensors cloned)."""
    import copy
    return {"state": {k: {n: (v.detach().clone() if torch.is_tensor(v) else copy.deepcopy(v)) for n, v in st.items()}
                      for k, st in sd["state"].items()},
            "param_groups": copy.deepcopy(sd["param_groups"])}

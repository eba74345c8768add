#!/usr/bin/env python
"""MMDiT rectified-flow training throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # product arm (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     # reference arm: CPU oracle port

Workload (configs[1] of BASELINE.json): MMDiT depth 12 / dim 768 / 12 heads, 256 px
(32x32x16 latent, patch 2 -> 256 image tokens + 154 text tokens), batch 64 per GPU, bf16,
synthetic latents and Gemma/CLIP-shaped text embeddings, random-init weights.
A step = noise -> forward -> velocity loss -> backward -> (grad all-reduce) -> clip -> AdamW.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))

CFG2 = dict(inCh=16, class_dim=768, patch_size=2, dim=768, hidden_scale=4.0, num_heads=12,
            attn_type="softmax_flash", MLP_type="swiglu", num_blocks=12, positional_encoding="RoPE2d")
LATENT, TEXT_TOKENS, BATCH = 32, 154, 64
METRIC, UNIT = "mmdit_train_images_per_sec_256px", "images/s"
# --config: the other BASELINE training configurations (parity-test shapes; NOT the bench line the
# metric is quoted on, which stays cfg2).  name -> (model overrides, latent side, batch per GPU, label)
CONFIGS = {
    "cfg2": (dict(), 32, 64, "MMDiT depth12/dim768/12 heads, 256px (32x32x16 latent, 256+154 tokens), "
                             "rectified-flow train step incl. clip+AdamW (BASELINE configs[1])"),
    "cfg3": (dict(dim=1536, num_heads=24, num_blocks=24), 32, 32,
             "MMDiT depth24/dim1536/24 heads, 256px (256+154 tokens), train step (BASELINE configs[2])"),
    "cfg4": (dict(dim=1536, num_heads=24, num_blocks=24), 64, 16,
             "MMDiT depth24/dim1536/24 heads, 512px (64x64x16 latent, 1024+154 tokens), train step (BASELINE configs[3])"),
}
WORKLOAD = CONFIGS["cfg2"][3]
REF_SAMPLE_BATCH = 2     # images per step of the CPU legs (reference arm and cpu_baseline): a bounded sample


def train_flops_per_image(cfg, N, M):
    """Algorithmic FLOPs: fwd = blocks + front/back-end GEMMs, train = 3 x fwd."""
    d, depth, C, p = cfg["dim"], cfg["num_blocks"], cfg["inCh"], cfg["patch_size"]
    T = N + M
    blocks = depth * (32 * T * d * d + 4 * T * T * d + 26 * d * d) - (26 * M * d * d + 8 * d * d)
    front = 2 * N * (C * p * p) * d + 2 * N * d * d + 2 * M * 2304 * d + 2 * 768 * d + 2 * d * d
    back = 4 * d * d + 2 * N * d * (C * p * p)
    attn = depth * 4 * T * T * d
    return 3.0 * (blocks + front + back), 3.0 * attn


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    hbm=p["hbm_gbs"], src="measured")
    except Exception:  # noqa: BLE001
        return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                mhz.append(float(r[0])); mx = float(r[1])
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mhz.sort()
        load = [m for m in mhz if mx and m > 0.3 * mx] or mhz
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx,
                "samples": len(mhz), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU oracle
def cpu_oracle_run(steps, warmup, sample_batch):
    """The reference algorithm (oracle/mmdit_oracle.py, pinned to the reference by golden vectors)
    on the host cores: fp32, all threads, `sample_batch` images per step of the cfg2 workload."""
    import torch
    from oracle import mmdit_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = dict(CFG2, attn_type="softmax", device="cpu")
    from src.models.diff_model import diff_model  # only for the state_dict schema (no compute)
    shapes = {k: tuple(v.shape) for k, v in diff_model(device="cpu", **CFG2).state_dict().items()}
    tr = O.TrainOracle(O.synth_state_dict(shapes), cfg)
    times = []
    for s in range(warmup + steps):
        b = O.synth_batch(sample_batch, CFG2["inCh"], LATENT, LATENT, TEXT_TOKENS, seed=1000 + s)
        t0 = time.perf_counter()
        tr.step(b)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return sample_batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    B = REF_SAMPLE_BATCH      # the same bounded sample as the product arm's cpu_baseline leg, whatever --steps is
    ips, sec = cpu_oracle_run(args.steps, args.warmup, B)
    cores = os.cpu_count() or 1
    sample = f"oracle port (fp32 torch CPU, {cores} threads), batch {B} per step of the cfg2 workload"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MMDiT depth12/dim768 256px rectified-flow train step (BASELINE configs[1])",
                   "batch_per_step": B, "tokens": 256 + TEXT_TOKENS,
                   "note": f"bounded sample: {B} images per step of the batch-{BATCH} workload (img/s is per-image, batch-size independent on the CPU)"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------- product arm
def _row_bytes(kind):
    """Algorithmic HBM bytes of the memory-bound kernels (SURVEY 8d, bf16 activations): (name, fn(args) -> bytes)."""
    return {
        "ln_modulate_fwd": lambda x, *a, **k: 4.0 * x.numel(),
        "ln_modulate_bwd": lambda dy, x, mean, rstd, scale, dres, *a, **k: (6.0 if dres is None else 8.0) * x.numel(),
        "gate_residual_fwd": lambda a_, *r, **k: 6.0 * a_.numel(),
        "gate_residual_ln_fwd": lambda a_, *r, **k: 8.0 * a_.numel(),      # read a, resid; write x', LN-mod(x')
        "gate_bwd": lambda dout, *r, **k: 6.0 * dout.numel(),              # read dout, a; write da (lines before round 2's last pass used 8)
        "qknorm_rope_fwd": lambda qkv, wq, wk, rope, d, *r, **k: 8.0 * qkv.shape[0] * d,
        "qknorm_rope_bwd": lambda dqk, qkv, wq, wk, rope, dqkv, dwq, dwk, d, *r, **k: 14.0 * qkv.shape[0] * d,   # dq fp32 4 + dk 2 + q,k 4 + dq,dk out 4
        "swiglu_bwd": lambda da, h12, *r, **k: 5.0 * h12.numel(),      # bf16: read da (h) + h12 (2h), write dh12 (2h)
    }[kind]


def instrument_kernels(trainer, batch, replays=3):
    """Per-kernel-class device time of one train step (forward + backward), free of host launch gaps:
    the step is captured ONCE into a CUDA graph on a single stream with an external-event record node
    before and after every GEMM / attention / row-kernel launch; the graph is replayed and the event
    pairs read back.  (Round 1 timed an eager step behind a `_sleep`; on a slow host the pairs timed
    launch gaps -- the driver's line showed GEMMs taking 3x the step.)  Serialised kernels on one
    stream: a pair brackets its kernel alone, and the classes can only sum to <= the replay time."""
    import torch
    from mmdit import ops, streams
    kinds = ["gemm", "attn_fwd", "attn_bwd", "ln_modulate_fwd", "ln_modulate_bwd", "gate_residual_fwd",
             "gate_residual_ln_fwd", "gate_bwd",
             "qknorm_rope_fwd", "qknorm_rope_bwd", "swiglu_bwd"]
    rec = {k: [] for k in kinds}
    orig = {k: getattr(ops, k) for k in kinds}

    def timed(kind, fn, work):
        def wrap(*a, **k):
            e0 = torch.cuda.Event(enable_timing=True, external=True)
            e1 = torch.cuda.Event(enable_timing=True, external=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            rec[kind].append((e0, e1, work(*a, **k)))
            return out
        return wrap

    def gemm_work(A, B, **k):
        M, K = (A.shape[1], A.shape[0]) if k.get("a_major") else A.shape
        N = B.shape[1] if k.get("b_major") else B.shape[0]
        return 2.0 * M * N * K

    work = {"gemm": gemm_work,
            "attn_fwd": lambda q, k, v, B, H, N, M, s, **kw: 4.0 * B * H * (N + M) ** 2 * 64,
            "attn_bwd": lambda q, k, v, o, l, do, dq, dk, dv, B, H, N, M, s: 10.0 * B * H * (N + M) ** 2 * 64}
    for k in kinds:
        setattr(ops, k, timed(k, orig[k], work.get(k) or _row_bytes(k)))
    dual = streams.ENABLED
    streams.ENABLED = False    # one stream: an event pair must bracket its kernel alone
    g = torch.cuda.CUDAGraph()
    empty = []                 # back-to-back event pairs: what a pair reads with nothing in between
    try:
        trainer._zero()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            if trainer.buckets is not None:
                trainer.buckets.reset()      # pins the CAPTURING stream as the one the exchange forks from
            trainer._fwd_bwd(batch)
            if trainer.buckets is not None:
                trainer.buckets.finish()
            for _ in range(16):
                a_ = torch.cuda.Event(enable_timing=True, external=True)
                b_ = torch.cuda.Event(enable_timing=True, external=True)
                a_.record(); b_.record()
                empty.append((a_, b_))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(replays):
            t0.record()
            g.replay()
            t1.record()
        torch.cuda.synchronize()
        replay_ms = t0.elapsed_time(t1)
    finally:
        for k in kinds:
            setattr(ops, k, orig[k])
        streams.ENABLED = dual
    pair_ms = sorted(a_.elapsed_time(b_) for a_, b_ in empty)[len(empty) // 2]
    out = {"replay_ms": replay_ms, "event_pair_overhead_us": 1e3 * pair_ms}
    for kind, items in rec.items():
        # a pair's reading includes the node-to-node hand-over that an empty pair also shows: subtract it
        ms = sum(max(0.0, e0.elapsed_time(e1) - pair_ms) for e0, e1, _ in items)
        out[kind] = dict(launches=len(items), ms=ms, work=sum(w for _, _, w in items))
    return out


def gpu_library_run(steps, warmup, batch, dev):
    """The reference's GPU path on this box ("the kernel to beat", SURVEY 8d / BASELINE.md section 4): the
    oracle's restatement of the reference modules under torch.autocast(bf16) -- cuBLAS GEMMs, ATen
    LayerNorm / RMSNorm / elementwise, flash_attn_func 2.8.3 as in Attention.py:293 (torch SDPA if
    flash-attn is not importable) -- with clip_grad_norm_ + torch.optim.AdamW(fused) as model_trainer.py
    :483-503.  Eager PyTorch, no activation checkpointing (the faster of the reference's two settings)."""
    import torch
    from oracle import mmdit_oracle as O
    from src.models.diff_model import diff_model
    try:
        import flash_attn  # noqa: F401
        O.ATTENTION_KERNEL = "flash"
    except Exception:  # noqa: BLE001
        O.ATTENTION_KERNEL = "sdpa"
    shapes = {k: tuple(v.shape) for k, v in diff_model(device="cpu", **CFG2).state_dict().items()}
    P = {k: v.to(dev).requires_grad_(not k.endswith("freqs")) for k, v in O.synth_state_dict(shapes).items()}
    params = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, eps=1e-8, weight_decay=0.01, fused=True)
    cfg = dict(CFG2, attn_type="softmax")
    bs = [{k: v.to(dev) for k, v in O.synth_batch(batch, CFG2["inCh"], LATENT, LATENT, TEXT_TOKENS, seed=1000 + i).items()}
          for i in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for s in range(warmup + steps):
            if s == warmup:
                torch.cuda.synchronize()
                e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss, _ = O.rf_loss(P, cfg, bs[s % 2])
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": batch,
                "last_loss": float(loss), "attention": O.ATTENTION_KERNEL,
                "what": "reference modules restated in eager PyTorch under autocast(bf16): cuBLAS + ATen + "
                        f"{'flash_attn_func 2.8.3' if O.ATTENTION_KERNEL == 'flash' else 'torch SDPA'}, "
                        "clip + torch fused AdamW, no checkpointing"}
    finally:
        O.ATTENTION_KERNEL = "eager"


def run_product(args):
    import torch
    import torch.distributed as dist
    from mmdit import _lib
    from mmdit.train import HostFeed, RFTrainer, host_batch
    from src.models.diff_model import diff_model

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mmdit_device_check(), "mmdit_device_check")

    torch.manual_seed(0)
    model = diff_model(device=dev, **CFG2)
    if world > 1:  # DDP ctor semantics: every replica starts from rank 0's weights
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    trainer = RFTrainer(model, world_size=world, use_graph=not args.no_graph,
                        peer=None if args.exchange == "peer" else False)
    exchange = ("none" if world == 1 else
                "peer-memory kernel per bucket, overlapped with backward, inside the step graph"
                if trainer.buckets.arena is not None else "NCCL all-reduce (torch.distributed)")
    C = CFG2["inCh"]
    nb = 4  # distinct host batches (3.3 MB each of text would be L2-resident; activations are not:
    #         one step touches > 20 GB of HBM, far beyond the 126 MB L2, so no explicit flush is needed)
    hbs = [host_batch(BATCH, C, LATENT, LATENT, TEXT_TOKENS, seed=1000 + rank * 9973 + i) for i in range(nb)]
    dbs = [trainer.to_device(hb) for hb in hbs]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step_fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def fresh(i):
        # device-resident batch i.  The model masks c / pooled in place (diff_model.py:281-287): under the
        # step graph that happens on the graph's static copy; eagerly it is idempotent for a fixed mask.
        return dbs[i % nb]

    # ---- device-resident arm ("value")
    for i in range(args.warmup):
        trainer.step(fresh(i))
    c0 = _lib.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed_loop(lambda i: trainer.step(fresh(i)), args.steps)
    clk = clocks.stop() if rank == 0 else None
    eager_launches = _lib.launch_count() - c0
    value = world * BATCH * args.steps / (ms / 1e3)

    # ---- end-to-end arm: pinned host batch -> H2D every step, loss -> host every step
    # (the H2D copy of batch i+1 is issued on a side stream while step i runs: mmdit.train.HostFeed)
    losses = []
    feed = HostFeed(dev)
    feed.submit(hbs[0])

    def e2e_step(i):
        batch = feed.take()
        feed.submit(hbs[(i + 1) % nb])      # next step's inputs start moving now
        loss = trainer.step(batch)
        losses.append(float(loss))          # D2H read of the step's result (sync, like :472-473)

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed_loop(e2e_step, args.steps)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)

    # ---- per-kernel-class roofline: one extra, untimed step captured on ONE stream with event nodes
    # around every GEMM / attention / row kernel (instrument_kernels).  Runs on EVERY rank: with data
    # parallelism the backward launches the gradient exchange.
    def close_reduce():
        if trainer.buckets is not None:
            trainer.buckets.finish()
    c1 = _lib.launch_count()
    trainer._zero(); trainer._fwd_bwd(fresh(1)); close_reduce()   # eager warm-up of the probe path
    launches_per_step = _lib.launch_count() - c1 + 3              # + the 3 optimizer kernels of a real step
    torch.cuda.synchronize()
    kern = instrument_kernels(trainer, fresh(0))
    # optimizer pass alone (30 B / parameter: p, g, m, v read; p, m, v, bf16 shadow written)
    n_param = sum(p.numel() for p in trainer.params)
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if trainer.fused_optimizer:
        saved_lr = trainer.opt.param_groups[0]["lr"]
        trainer.opt.param_groups[0]["lr"] = 0.0      # measured, not trained
        trainer.opt.step()
        o0.record(); trainer.opt.step(); o1.record()
        torch.cuda.synchronize()
        trainer.opt.param_groups[0]["lr"] = saved_lr
        trainer.opt.sync_lr()
        kern["adamw"] = dict(launches=1, ms=o0.elapsed_time(o1), work=30.0 * n_param)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    f_img, f_attn_img = train_flops_per_image(CFG2, (LATENT // 2) ** 2, TEXT_TOKENS)
    step_ms = ms / args.steps
    g = kern["gemm"]
    gemm_tf = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0

    def cls(kind, unit_scale, peak):
        k = kern[kind]
        if not k["launches"] or not k["ms"]:
            return None
        ach = k["work"] / (k["ms"] * 1e-3) / unit_scale
        return {"achieved": ach, "frac": ach / peak, "ms_per_step": k["ms"], "launches_per_step": k["launches"],
                "avg_launch_us": 1e3 * k["ms"] / k["launches"]}
    tensor_classes = {k: cls(k, 1e12, pk["tf_sustained"]) for k in ("attn_fwd", "attn_bwd")}
    hbm_classes = {k: cls(k, 1e9, pk["hbm"]) for k in kern
                   if k not in ("gemm", "attn_fwd", "attn_bwd", "replay_ms", "event_pair_overhead_us")}
    probe_sum_ms = sum(v["ms"] for k, v in kern.items() if isinstance(v, dict) and k != "adamw")
    # the probe step is serialised on one stream: instrumented classes can only sum to <= its replay time
    probe_ok = probe_sum_ms <= kern["replay_ms"] * 1.02 and g["ms"] <= step_ms
    traffic, traffic_src = None, None
    if args.config == "cfg2":
        # NOT measured in this run (DRAM counters need ncu): the committed capture of this same step
        try:
            with open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")) as f:
                traffic = json.load(f)["gemm_tcgen05_kernel"]["dram_bytes_per_launch"]
            traffic_src = ("constant from profiles/r02_kernel_metrics.json (ncu dram__bytes_read+write, mean over "
                           "the cfg2 step's GEMM launches); not re-measured by bench.py")
        except Exception:  # noqa: BLE001
            pass
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel",
        "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tf / pk["tf_sustained"], "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
        "launches_per_step": g["launches"], "avg_launch_us": 1e3 * g["ms"] / max(1, g["launches"]),
        "share_of_step": g["ms"] / step_ms, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_flops_per_step": g["work"],
        "method": "external CUDA-event nodes around every launch of a single-stream graph capture of the step",
        "probe": {"replay_ms": kern["replay_ms"], "instrumented_ms": probe_sum_ms, "consistent": bool(probe_ok),
                  "event_pair_overhead_us": kern["event_pair_overhead_us"]},
        "tensor_classes": tensor_classes,
        "hbm_classes": {"peak": pk["hbm"], "unit": "GB/s", **hbm_classes},
        "attention": {"achieved": (kern["attn_fwd"]["work"] + kern["attn_bwd"]["work"]) /
                      max(1e-9, (kern["attn_fwd"]["ms"] + kern["attn_bwd"]["ms"]) * 1e-3) / 1e12,
                      "unit": "TFLOP/s", "ms_per_step": kern["attn_fwd"]["ms"] + kern["attn_bwd"]["ms"],
                      "share_of_step": (kern["attn_fwd"]["ms"] + kern["attn_bwd"]["ms"]) / step_ms},
        "step_model_flops_utilisation": f_img * BATCH / (step_ms * 1e-3) / 1e12 / pk["tf_sustained"],
    }
    roofline["attention"]["frac"] = roofline["attention"]["achieved"] / pk["tf_sustained"]
    if not probe_ok:
        roofline["invalid"] = "per-class times do not fit inside the step: probe rejected"

    # ---- the reference's GPU path on this box (library kernels), same workload, rank 0 at N=1 only
    lib = None
    used_graph = bool(trainer.use_graph)
    if world == 1 and args.config == "cfg2" and not args.no_gpu_library_baseline:
        try:
            del trainer, model
            torch.cuda.empty_cache()
            lib = gpu_library_run(3, 2, BATCH, dev)
        except Exception as e:  # noqa: BLE001
            lib = {"value": None, "error": repr(e)[:300]}

    # ---- CPU baseline on this box's host cores (bounded sample: 1 warm-up + 2 steps of batch 2)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            ips, sec = cpu_oracle_run(2, 1, REF_SAMPLE_BATCH)
            cpu = {"value": ips, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"oracle port fp32, 2 steps of batch 2 of the same cfg2 workload ({sec:.2f} s/step)"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"failed: {e}"}

    h2d = sum(v.numel() * v.element_size() for v in hbs[0].values())
    finite = all(l == l and abs(l) != float("inf") for l in losses)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": world * BATCH, "batch_per_gpu": BATCH, "parallelism": f"dp{world}",
                   "cuda_graph": used_graph, "gradient_exchange": exchange,
                   "two_stream_blocks": bool(__import__("mmdit.streams").streams.ENABLED),
                   "l2": "per-step working set (>20 GB) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": losses[-1]},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "launch_calls_in_timed_region": eager_launches,
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "gpu_library_baseline": lib,
        "tflops_per_step": f_img * BATCH / 1e12,
        **({} if finite else {"invalid": "non-finite loss in the timed region"}),
    }))
    if world > 1:
        dist.destroy_process_group()
    if not finite:
        sys.exit("bench.py: non-finite loss -- the throughput above is not a valid measurement")


def select_config(name, batch):
    """Point the module-level workload constants at one of CONFIGS."""
    global CFG2, LATENT, BATCH, WORKLOAD, METRIC
    over, LATENT, b, WORKLOAD = CONFIGS[name]
    CFG2 = dict(CFG2, **over)
    BATCH = batch or b
    if name != "cfg2":
        METRIC = f"mmdit_train_images_per_sec_{LATENT * 8}px_{name}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step into a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-library-baseline", action="store_true",
                    help="skip the eager-PyTorch (cuBLAS + flash-attn) run of the same workload on the GPU")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="data-parallel gradient exchange: our peer-memory kernel (default) or NCCL")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    args = ap.parse_args()
    select_config(args.config, args.batch)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()

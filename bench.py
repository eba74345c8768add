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
    B = 2 if (args.steps + args.warmup) <= 24 else 1
    ips, sec = cpu_oracle_run(args.steps, args.warmup, B)
    cores = os.cpu_count() or 1
    sample = f"oracle port (fp32 torch CPU, {cores} threads), batch {B} per step of the cfg2 workload"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MMDiT depth12/dim768 256px rectified-flow train step (BASELINE configs[1])",
                   "batch_per_step": B, "tokens": 256 + TEXT_TOKENS},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------- product arm
def instrument_kernels(trainer, batch):
    """One extra, untimed step with CUDA events around every GEMM / attention launch (on the
    launching stream) -> per-class device time and algorithmic work for the roofline object."""
    import torch
    from mmdit import ops
    rec = {"gemm": [], "attn_fwd": [], "attn_bwd": []}
    orig = (ops.gemm, ops.attn_fwd, ops.attn_bwd)

    def timed(kind, fn, work):
        def wrap(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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

    ops.gemm = timed("gemm", orig[0], gemm_work)
    ops.attn_fwd = timed("attn_fwd", orig[1], lambda q, k, v, B, H, N, M, s, **kw: 4.0 * B * H * (N + M) ** 2 * 64)
    ops.attn_bwd = timed("attn_bwd", orig[2],
                         lambda q, k, v, o, l, do, dq, dk, dv, B, H, N, M, s: 10.0 * B * H * (N + M) ** 2 * 64)
    import mmdit.functional as Fn
    from mmdit import streams
    dual = streams.ENABLED
    streams.ENABLED = False    # one stream: an event pair must time its kernel alone, not a neighbour too
    try:
        trainer._zero()
        torch.cuda._sleep(int(4e8))   # keep the GPU busy while the host enqueues: events then time kernels, not launch gaps
        trainer._fwd_bwd(batch)
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.attn_fwd, ops.attn_bwd = orig
        streams.ENABLED = dual
    out = {}
    for kind, items in rec.items():
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in items)
        out[kind] = dict(launches=len(items), ms=ms, flops=sum(w for _, _, w in items))
    del Fn
    return out


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

    def fresh(i):  # the model masks c / pooled in place, so every step gets a pristine copy
        b = dbs[i % nb]
        return {k: (v.clone() if k in ("c", "pooled") else v) for k, v in b.items()}

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

    # ---- per-kernel-class roofline from an instrumented, untimed step (eager, same shapes).
    # Runs on EVERY rank: with data parallelism the backward launches gradient all-reduces.
    # (uses the trainer's eager building blocks; no optimizer step, nothing is replayed afterwards)
    def close_reduce():
        if trainer.buckets is not None:
            trainer.buckets.finish()
    trainer._zero(); trainer._fwd_bwd(fresh(1)); close_reduce()   # eager warm-up of the probe path
    c1 = _lib.launch_count()
    kern = instrument_kernels(trainer, fresh(0))
    launches_per_step = _lib.launch_count() - c1 + 3              # + the 3 optimizer kernels of a real step
    close_reduce()
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    f_img, f_attn_img = train_flops_per_image(CFG2, (LATENT // 2) ** 2, TEXT_TOKENS)
    g = kern["gemm"]
    gemm_tf = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0
    att_ms = kern["attn_fwd"]["ms"] + kern["attn_bwd"]["ms"]
    att_tf = (kern["attn_fwd"]["flops"] + kern["attn_bwd"]["flops"]) / (att_ms * 1e-3) / 1e12 if att_ms else 0.0
    step_ms = ms / args.steps
    traffic = None   # dram__bytes_read+write per GEMM launch, from the committed ncu capture of this step
    try:
        with open(os.path.join(ROOT, "profiles", "r01_kernel_metrics_v6.json")) as f:
            traffic = json.load(f)["gemm_tcgen05_kernel"]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel",
        "achieved": gemm_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tf / pk["tf_sustained"], "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
        "launches_per_step": g["launches"], "avg_launch_us": 1e3 * g["ms"] / max(1, g["launches"]),
        "share_of_step": g["ms"] / step_ms, "traffic": traffic,
        "traffic_source": "profiles/r01_kernel_metrics_v6.json (ncu, mean over the step's 340 GEMM launches)",
        "algorithmic_flops_per_step": g["flops"],
        "attention": {"achieved": att_tf, "unit": "TFLOP/s", "frac": att_tf / pk["tf_sustained"],
                      "ms_per_step": att_ms, "share_of_step": att_ms / step_ms},
        "step_model_flops_utilisation": f_img * BATCH / (step_ms * 1e-3) / 1e12 / pk["tf_sustained"],
    }

    # ---- CPU baseline on this box's host cores (bounded sample: 1 warm-up + 2 steps of batch 2)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            ips, sec = cpu_oracle_run(2, 1, 2)
            cpu = {"value": ips, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"oracle port fp32, 2 steps of batch 2 of the same cfg2 workload ({sec:.2f} s/step)"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"failed: {e}"}

    h2d = trainer.h2d_bytes(hbs[0])
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": world * BATCH, "batch_per_gpu": BATCH, "parallelism": f"dp{world}",
                   "cuda_graph": bool(trainer.use_graph), "gradient_exchange": exchange,
                   "two_stream_blocks": bool(__import__("mmdit.streams").streams.ENABLED),
                   "l2": "per-step working set (>20 GB) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": losses[-1]},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "launch_calls_in_timed_region": eager_launches,
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "tflops_per_step": f_img * BATCH / 1e12,
    }))
    if world > 1:
        dist.destroy_process_group()


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

"""Multi-GPU check of the peer-memory gradient exchange (run under torchrun, one rank per GPU):
  1. arena all-reduce-mean vs NCCL all_reduce on random buckets (eager), bit-identical across ranks
  2. the same launches captured in a CUDA graph and replayed with fresh data (epoch logic)
  3. a small MMDiT trained for a few steps: peer exchange inside ONE graph vs NCCL eager exchange
"""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
from mmdit.comm import PeerArena
from mmdit.train import RFTrainer, host_batch
from src.models.diff_model import diff_model

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def same_everywhere(t):
    ref = t.clone()
    dist.broadcast(ref, 0)
    return bool((ref == t).all())


# ---- 1. eager
N = 6_000_000
arena = PeerArena(N, dev)
ranges = [(0, 1_000_000), (1_000_000, 3_000_064), (4_000_064, 1_999_936 - 64)]
g = torch.Generator(device=dev).manual_seed(100 + rank)
for it in range(3):
    arena.flat.copy_(torch.randn(arena.n, device=dev, generator=g))
    ref = arena.flat.clone()
    dist.all_reduce(ref)
    ref /= world
    torch.cuda.synchronize(); dist.barrier()
    for off, n in ranges:
        arena.all_reduce_mean(off, n)
    torch.cuda.synchronize()
    covered = torch.cat([arena.flat[o:o + n] for o, n in ranges]); refc = torch.cat([ref[o:o + n] for o, n in ranges])
    err = float((covered - refc).abs().max())
    same = same_everywhere(covered)
    say(f"[eager {it}] max |peer - nccl| = {err:.3e}, identical on all ranks: {same}")
    ok &= err < 1e-5 and same
    dist.barrier()

# ---- 2. captured + replayed
side = torch.cuda.Stream()
static = torch.zeros(arena.n, device=dev)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    arena.flat.copy_(static)
    main = torch.cuda.current_stream()
    for off, n in ranges:
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        arena.all_reduce_mean(off, n, stream=side)
        arena.flat[:16].add_(0.0)          # main-stream work next to the exchange
    ev = torch.cuda.Event(); ev.record(side); main.wait_event(ev)
for it in range(4):
    static.copy_(torch.randn(arena.n, device=dev, generator=g))
    ref = static.clone(); dist.all_reduce(ref); ref /= world
    torch.cuda.synchronize(); dist.barrier()
    graph.replay()
    torch.cuda.synchronize()
    covered = torch.cat([arena.flat[o:o + n] for o, n in ranges]); refc = torch.cat([ref[o:o + n] for o, n in ranges])
    err = float((covered - refc).abs().max())
    same = same_everywhere(covered)
    say(f"[graph {it}] max |peer - nccl| = {err:.3e}, identical on all ranks: {same}")
    ok &= err < 1e-5 and same
    dist.barrier()

# bandwidth of one big bucket (eager, back to back)
big = (0, 5_999_936)
for _ in range(3):
    arena.all_reduce_mean(*big)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    arena.all_reduce_mean(*big)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
say(f"[bw] {big[1] * 4 / 1e6:.1f} MB bucket: {ms * 1e3:.1f} us -> algbw {big[1] * 4 / ms / 1e6:.1f} GB/s "
    f"(busbw {2 * (world - 1) / world * big[1] * 4 / ms / 1e6:.1f} GB/s)")
ref = torch.randn(big[1], device=dev)
for _ in range(3):
    dist.all_reduce(ref)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    dist.all_reduce(ref)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
say(f"[bw] NCCL all_reduce same size: {ms * 1e3:.1f} us -> algbw {big[1] * 4 / ms / 1e6:.1f} GB/s")
arena.close()

# ---- 3. trainer: peer exchange in one graph vs NCCL eager
cfg = dict(inCh=4, class_dim=768, patch_size=2, dim=256, hidden_scale=4.0, num_heads=4,
           attn_type="softmax_flash", MLP_type="swiglu", num_blocks=2, positional_encoding="RoPE2d")
losses = {}
for mode in ("nccl_eager", "peer_eager", "nccl_graph", "peer_graph"):
    torch.manual_seed(0)
    model = diff_model(device=dev, **cfg)
    for p in model.parameters():
        dist.broadcast(p.data, 0)
    tr = RFTrainer(model, world_size=world, use_graph=mode.endswith("graph"), peer=mode.startswith("peer"))
    ls = []
    for s in range(6):
        hb = host_batch(4, 4, 32, 32, 154, seed=500 + s * world + rank)
        torch.manual_seed(1234 + s)            # same epsilon stream in every mode
        ls.append(float(tr.step(tr.to_device(hb))))
    torch.cuda.synchronize()
    if not mode.endswith("graph"):
        say(f"    gradient elements written straight into their bucket slot: {tr.buckets.direct_elems}, copied in: {tr.buckets.copied_elems}")
    w = model.blocks[0].MLP_x.MLP.w12.weight.detach()
    losses[mode] = (ls, float(w.double().abs().sum()), same_everywhere(w))
    say(f"[train {mode}] losses {[round(x, 5) for x in ls]} |w12| {losses[mode][1]:.6f} replicas identical: {losses[mode][2]}")
    if tr.buckets is not None and tr.buckets.arena is not None:
        tr.buckets.arena.close()
    dist.barrier()
for mode, base in (("peer_eager", "nccl_eager"), ("peer_graph", "nccl_graph")):
    d = max(abs(a - b) for a, b in zip(losses[base][0], losses[mode][0]))
    say(f"[train] max |loss({mode}) - loss({base})| = {d:.2e}")
    ok &= d < 2e-3 and losses[mode][2]
say("DDP_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()

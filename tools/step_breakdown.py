"""Per-shape GEMM time inside one eager cfg2 training step (CUDA events on the launching stream,
GPU pre-loaded so that events time kernels, not host launch gaps)."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
import bench
from mmdit import ops
from mmdit.train import RFTrainer, host_batch
from src.models.diff_model import diff_model

dev = torch.device("cuda"); torch.manual_seed(0)
model = diff_model(device=dev, **bench.CFG2)
tr = RFTrainer(model, use_graph=False)
hb = host_batch(bench.BATCH, 16, 32, 32, 154, seed=1)
for _ in range(3):
    tr.step(tr.to_device(hb))
rec = []
orig = ops.gemm
def timed(A, B, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(A, B, **k); e1.record()
    M, K = (A.shape[1], A.shape[0]) if k.get("a_major") else A.shape
    N = B.shape[1] if k.get("b_major") else B.shape[0]
    rec.append(((M, N, K, int(k.get("a_major", 0)), int(k.get("b_major", 0)), int(k.get("epilogue", 0)),
                 "f32" if out.dtype == torch.float32 else "bf16", k.get("bias") is not None), e0, e1))
    return out
ops.gemm = timed
b = tr.to_device(hb)
torch.cuda.synchronize()
torch.cuda._sleep(int(1.5e9))
tr.step(b)
torch.cuda.synchronize()
ops.gemm = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in rec:
    agg[key][0] += 1; agg[key][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"{len(rec)} GEMM launches, {tot:.2f} ms")
print(f"{'M':>6} {'N':>6} {'K':>6} aM bM epi out  bias  count     ms   us/call   TFLOP/s")
for key, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, am, bm, epi, dt, bias = key
    print(f"{M:6d} {N:6d} {K:6d} {am:2d} {bm:2d} {epi:3d} {dt:4s} {int(bias):4d} {c:6d} {ms:7.3f} {1e3*ms/c:8.1f} {2.0*M*N*K*c/ms/1e9:9.1f}")

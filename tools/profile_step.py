"""One cfg2 training step between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-3-from-scratch_b200"))
import bench  # noqa: E402
from mmdit.train import RFTrainer, host_batch  # noqa: E402
from src.models.diff_model import diff_model  # noqa: E402

B = int(os.environ.get("PROFILE_BATCH", bench.BATCH))
dev = torch.device("cuda")
torch.manual_seed(0)
model = diff_model(device=dev, **bench.CFG2)
tr = RFTrainer(model, use_graph=False)
hb = host_batch(B, 16, 32, 32, 154, seed=1)
for _ in range(2):
    tr.step(tr.to_device(hb))
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(tr.to_device(hb))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")

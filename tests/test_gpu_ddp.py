"""GPU (-m gpu), needs >= 2 devices: the data-parallel gradient exchange (csrc/comm.cu, mmdit/comm.py,
GradBuckets) on real NVLink peers -- replaces the DDP reducer of model_trainer.py:224.  Runs
tools/ddp_check.py under torchrun: the peer-memory all-reduce-mean must equal NCCL's all_reduce / W on
random buckets (eager and replayed from a CUDA graph, bit-identical replicas), and a small MMDiT trained
with the peer exchange inside ONE step graph must follow the NCCL-exchange trajectory."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_exchange_matches_nccl_on_all_visible_gpus():
    n = torch.cuda.device_count()
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    port = 29800 + os.getpid() % 100
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "DDP_CHECK PASS" in r.stdout

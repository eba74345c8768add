"""Two-stream schedule of the dual-stream MMDiT block (Transformer_Block_Dual.py:56-78).

The image branch and the text branch of a block only meet at the joint attention.  The image
branch stays on the caller's stream and the text branch runs on one side stream per device;
the three sync points per block (modulation ready -> text branch, text QKV ready -> attention,
attention output ready -> text branch) are events, so a captured training step becomes a CUDA
graph with two parallel chains.  The hardware then starts the text-branch GEMM on the SMs that
the image-branch GEMM's last (partial) wave leaves idle, and runs memory-bound row kernels of
one branch next to tensor-core kernels of the other.

Used only while gradients are recorded: every tensor that crosses the streams in the forward
is then saved for the backward (no early free that the caching allocator could recycle on the
other stream); autograd runs each backward node on its forward stream and orders / records the
gradients that cross.  MMDIT_DUAL_STREAM=0 turns it off.
"""
import os

import torch

ENABLED = os.environ.get("MMDIT_DUAL_STREAM", "1") == "1"
_side = {}


def side(device):
    """The text-branch stream of `device` (created on first use)."""
    key = torch.device(device).index
    if key is None:
        key = torch.cuda.current_device()
    s = _side.get(key)
    if s is None:
        s = _side[key] = torch.cuda.Stream(device=key)
    return s


# Under no_grad (sampling) nothing is saved for a backward, so a tensor that crosses the streams
# could be freed -- and recycled by the caching allocator on its home stream -- while the other
# stream still reads it.  With MMDIT_DUAL_STREAM_INFER=1 the blocks park those tensors in
# `infer_keepalive` until the end-of-forward join (diff_model.forward clears it).  Off by default:
# not yet timed / validated on hardware.
INFER = os.environ.get("MMDIT_DUAL_STREAM_INFER", "0") == "1"
infer_keepalive = []


def active(t):
    return ENABLED and t.is_cuda and (torch.is_grad_enabled() or INFER)


def hold(*tensors):
    """Keep tensors that cross the streams alive until the forward's final join (no_grad only)."""
    if not torch.is_grad_enabled():
        infer_keepalive.extend(tensors)


# ---- weight gradients off the critical path ------------------------------------------------
# In the backward only the data gradients form a chain; every weight gradient is a leaf that
# nobody reads before the optimizer.  The trainer (mmdit.train.RFTrainer) sets `async_wgrad`
# around its backward: wgrad GEMMs then run on a third stream (forked after their operands are
# ready), the data-gradient chain never waits for them, and the trainer joins the stream before
# the gradients are consumed (all-reduce / clip / AdamW).  Operands are kept alive in `keepalive`
# until the next step so that the caching allocator cannot recycle them under the wgrad stream.
# Never on for callers who run backward themselves (they would have to join the stream).
# Measured on B200 (cfg2, same box, 20 steps): 31.35 / 31.75 ms off vs 31.42 / 31.11 ms on -- inside
# the noise, because the persistent GEMMs already fill the machine; OFF by default (MMDIT_ASYNC_WGRAD=1).
ASYNC_WGRAD = os.environ.get("MMDIT_ASYNC_WGRAD", "0") == "1"
async_wgrad = False
keepalive = []
_wgrad = {}


def wgrad(device):
    key = torch.device(device).index
    if key is None:
        key = torch.cuda.current_device()
    s = _wgrad.get(key)
    if s is None:
        s = _wgrad[key] = torch.cuda.Stream(device=key)
    return s


def join_wgrad(device):
    """Make the current stream wait for every weight gradient issued so far."""
    if _wgrad:
        torch.cuda.current_stream().wait_stream(wgrad(device))

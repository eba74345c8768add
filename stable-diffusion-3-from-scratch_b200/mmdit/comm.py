"""Peer-memory gradient exchange for data parallelism (model_trainer.py:224 -- the DDP reducer).

Each rank allocates ONE IPC-exportable arena (fp32 gradients of every bucket + a signal pad)
through the C-ABI, the ranks swap CUDA IPC handles over torch.distributed, and every bucket is
then reduced by a single captured-in-graph kernel (csrc/comm.cu: reduce-scatter + all-gather +
mean over NVLink loads/stores).  torch.distributed is only the rendezvous here, not the data path.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib

PAD_BYTES = 256      # signal pad: uint32[2][8] flags + uint32[2] kernel state, padded
ALIGN = 64           # buckets start on 256-byte boundaries (64 floats)


class CommStruct(C.Structure):
    """Mirror of `mmdit_comm` (include/mmdit_b200.h)."""

    _fields_ = [("buf", C.c_void_p * 8), ("flag", C.c_void_p * 8), ("state", C.c_void_p),
                ("world", C.c_int32), ("rank", C.c_int32)]


class _DeviceSpan:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2, "strides": None}


class PeerArena:
    """fp32 arena of `n_floats` on this rank, mapped into every peer of `group`."""

    def __init__(self, n_floats, device, group=None):
        L = _lib.lib()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world not in (1, 2, 4, 8):
            raise ValueError(f"peer-memory all-reduce supports 1, 2, 4 or 8 ranks, not {self.world}")
        self.n = (n_floats + ALIGN - 1) // ALIGN * ALIGN
        self.bytes = self.n * 4 + 2 * PAD_BYTES
        base = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.mmdit_comm_alloc(C.byref(base), self.bytes), "mmdit_comm_alloc")
            self.base = base.value
            hb = L.mmdit_comm_handle_bytes()
            handle = (C.c_ubyte * hb)()
            _lib.check(L.mmdit_comm_export(self.base, handle), "mmdit_comm_export")
            mine = bytes(handle)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.peer_bases, self._opened = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_bases.append(self.base)
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * hb).from_buffer_copy(h)
                _lib.check(L.mmdit_comm_import(buf, C.byref(p)), f"mmdit_comm_import(rank {r})")
                self.peer_bases.append(p.value)
                self._opened.append(p.value)
        self.comm = CommStruct()
        for r, b in enumerate(self.peer_bases):
            self.comm.buf[r] = b
            self.comm.flag[r] = b + self.n * 4
        self.comm.state = self.base + self.n * 4 + PAD_BYTES
        self.comm.world, self.comm.rank = self.world, self.rank
        self.flat = torch.as_tensor(_DeviceSpan(self.base, self.n), device=device)
        dist.barrier(group=group)   # every rank has mapped every arena before the first kernel

    def all_reduce_mean(self, offset, n, stream=None, ctas=0):
        """arena[offset:offset+n] <- mean over ranks (one kernel on `stream`, graph capturable)."""
        ctas = ctas or int(os.environ.get("MMDIT_COMM_CTAS", "0"))
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mmdit_allreduce_mean_f32(C.byref(self.comm), offset, n, ctas, s),
                   "mmdit_allreduce_mean_f32")

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize()
        for p in self._opened:
            L.mmdit_comm_close(p)
        self._opened = []
        self.flat = None
        dist.barrier(group=self.group)   # nobody frees an arena a peer still has mapped
        if self.base:
            L.mmdit_comm_free(self.base)
            self.base = 0

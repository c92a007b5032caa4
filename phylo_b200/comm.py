"""The collectives particle sharding needs (include/vcsmc_b200.h, ``vcsmc_comm_fn``) over ``torch.distributed``.

torch is plumbing here: the data path between GPUs -- reading a remote ancestor's forest row, copying missing nodes out
of the owner's pool -- is done by the library's own kernels through peer pointers (CUDA IPC).  What is asked of
``torch.distributed`` is, per rank event, ONE all-gather of the step record and one barrier.

Backends:
  * ``nccl``: every call is stream-ordered on the current CUDA stream (NVLink / NVSwitch on a B200 box);
  * ``gloo``: used by the tests to run several ranks on ONE GPU (NCCL refuses two ranks per device); buffers are staged
    through host memory with a stream synchronisation on both sides -- correct, slow, never used for a benchmark.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


def _sync(t: torch.Tensor) -> None:
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()


class Comm:
    def __init__(self, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = str(dist.get_backend(group))
        self._one: Optional[torch.Tensor] = None

    @property
    def on_stream(self) -> bool:
        return "nccl" in self.backend

    # -- the three operations of vcsmc_comm_fn ----------------------------------------------------
    def all_gather_inplace(self, buf: torch.Tensor, chunk_bytes: int) -> None:
        """``buf`` (uint8, world * chunk_bytes): chunk g is rank g's contribution; filled in place on every rank."""
        mine = buf[self.rank * chunk_bytes:(self.rank + 1) * chunk_bytes]
        if self.on_stream:
            dist.all_gather_into_tensor(buf, mine.clone(), group=self.group)
            return
        _sync(buf)
        host = mine.cpu()
        parts = [torch.empty_like(host) for _ in range(self.world)]
        dist.all_gather(parts, host, group=self.group)
        buf.copy_(torch.cat(parts).to(buf.device))
        _sync(buf)

    def barrier(self, device) -> None:
        """Every rank's earlier work on the current stream is complete before any rank's later work starts."""
        if self.on_stream:
            if self._one is None:
                self._one = torch.zeros(1, dtype=torch.float32, device=device)
            dist.all_reduce(self._one, group=self.group)
            return
        if torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()
        dist.barrier(group=self.group)

    def all_reduce(self, t: torch.Tensor) -> None:
        if self.on_stream or not t.is_cuda:
            dist.all_reduce(t, group=self.group)
            return
        _sync(t)
        host = t.cpu()
        dist.all_reduce(host, group=self.group)
        t.copy_(host.to(t.device))
        _sync(t)

    # -- host-side agreement ------------------------------------------------------------------------
    def min_int(self, value: int) -> int:
        """The minimum of a host integer over the ranks (every rank must size its workspace identically)."""
        vals: List[int] = [0] * self.world
        dist.all_gather_object(vals, int(value), group=self.group)
        return min(vals)

    def exchange(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out


class PeerMap:
    """Every rank's workspace mapped into this process (CUDA IPC): the peer pointers vcsmc_sweep_set_comm takes."""

    def __init__(self, comm: Comm, workspace: torch.Tensor):
        lib = _lib.load()
        self._lib = lib
        self._opened: List[int] = []
        handle = (C.c_char * 64)()
        offset = C.c_int64(0)
        _lib.check(lib.vcsmc_ipc_export(workspace.data_ptr(), handle, C.byref(offset)))
        infos = comm.exchange((bytes(handle), int(offset.value), int(workspace.numel())))
        sizes = {i[2] for i in infos}
        if len(sizes) != 1:
            raise RuntimeError("ranks allocated workspaces of different sizes: %s" % sorted(sizes))
        self.ptrs: List[int] = []
        for g, (h, off, _n) in enumerate(infos):
            if g == comm.rank:
                self.ptrs.append(workspace.data_ptr())
                continue
            base = C.c_void_p()
            hb = (C.c_char * 64).from_buffer_copy(h)
            _lib.check(lib.vcsmc_ipc_open(hb, C.byref(base)))
            self._opened.append(base.value)
            self.ptrs.append(base.value + off)
        comm.barrier(workspace.device)   # nobody proceeds (or frees) before every peer has mapped every workspace

    def as_array(self):
        return (C.c_void_p * len(self.ptrs))(*self.ptrs)

    def close(self):
        opened, self._opened = self._opened, []
        for b in opened:
            self._lib.vcsmc_ipc_close(b)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

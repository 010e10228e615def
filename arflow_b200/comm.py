"""Peer-memory gradient all-reduce for data-parallel training (one process per GPU, one node).

Host side of csrc/comm.cu: every rank allocates its flat gradient buffer and a flag area with the library's own
allocator (CUDA-IPC exportable), the 64-byte handles travel through torch.distributed, and `all_reduce_` is then a
single kernel launch on the current stream — capturable in a CUDA graph, forkable onto a side stream while backward is
still running.  Replaces, for the benchmark driver, nn.DataParallel's gradient gather of the reference
(trainer/base_trainer.py:75,131-147).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class _DevBuffer:
    """A raw device allocation presented to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nfloats):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerAllReduce:
    """flat fp32 buffer of `numel` elements (rounded up to 4) shared for peer access among the ranks of `group`."""

    def __init__(self, numel, device, group=None, ctas=32):
        if not dist.is_initialized():
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("PeerAllReduce: at most 8 ranks (one node)")
        self.device = torch.device(device)
        self.numel = (int(numel) + 3) // 4 * 4
        self.ctas = int(ctas)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            self._data = ctypes.c_void_p()
            self._flags = ctypes.c_void_p()
            _lib.check(lib.arf_comm_alloc(ctypes.byref(self._data), self.numel * 4), "arf_comm_alloc")
            _lib.check(lib.arf_comm_alloc(ctypes.byref(self._flags), lib.arf_comm_flag_bytes()), "arf_comm_alloc")
            hd, hf = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
            _lib.check(lib.arf_comm_ipc_get(self._data, hd), "arf_comm_ipc_get")
            _lib.check(lib.arf_comm_ipc_get(self._flags, hf), "arf_comm_ipc_get")
            mine = (hd.raw, hf.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=self.group)
            self._peer_data = (ctypes.c_void_p * self.world)()
            self._peer_flags = (ctypes.c_void_p * self.world)()
            self._opened = []
            for p, (pd, pf) in enumerate(handles):
                if p == self.rank:
                    self._peer_data[p], self._peer_flags[p] = self._data.value, self._flags.value
                    continue
                a, b = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.check(lib.arf_comm_ipc_open(pd, ctypes.byref(a)), "arf_comm_ipc_open")
                _lib.check(lib.arf_comm_ipc_open(pf, ctypes.byref(b)), "arf_comm_ipc_open")
                self._peer_data[p], self._peer_flags[p] = a.value, b.value
                self._opened += [a, b]
            self._holder = _DevBuffer(self._data.value, self.numel)
            self.buffer = torch.as_tensor(self._holder, device=self.device)
        dist.barrier(group=self.group)     # every peer has mapped every buffer before the first kernel

    def all_reduce_(self, start=0, end=None, average=True, ctas=None):
        """In-place sum (or mean) of buffer[start:end] over the ranks; start and end multiples of 4.  ctas: CTAs of this
        launch (default: the constructor's) — few while other kernels should keep the SMs, up to 64 when nothing else
        runs; every rank must pass the same value."""
        end = self.numel if end is None else int(end)
        scale = 1.0 / self.world if average else 1.0
        _lib.call("arf_allreduce_f32", self._peer_data, self._peer_flags, self.rank, self.world, int(start),
                  end - int(start), scale, int(ctas or self.ctas), _lib.stream_ptr())

    def check(self):
        """Synchronise and raise if a barrier timed out (a peer never launched its kernel)."""
        _lib.call("arf_comm_error", self._flags, _lib.stream_ptr())

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for h in self._opened:
            lib.arf_comm_ipc_close(h)
        self._opened = []
        self.buffer = None
        self._holder = None
        dist.barrier(group=self.group)
        lib.arf_comm_free(self._data)
        lib.arf_comm_free(self._flags)

"""Multi-GPU plumbing: one process per GPU, cells (columns of Y) sharded contiguously
over ranks.  The data path's collectives are NCCL calls inside
libscgbm.so; the host side only has to (a) pick each rank's column range, (b) hand
every rank the same NCCL unique id, (c) agree on scalars such as nbatch.  Any
broadcast transport works for (b) -- bench.py and the tests use torch.distributed (gloo) there
(tests/dist_util.py); (c) runs over the library's own NCCL all-reduce.  This package imports no torch."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def shard_range(J: int, rank: int, nranks: int):
    """Contiguous, balanced split of J cells: ranks < J % nranks get one more."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    base, rem = divmod(J, nranks)
    j0 = rank * base + min(rank, rem)
    j1 = j0 + base + (1 if rank < rem else 0)
    return j0, j1


class Comm:
    """One NCCL rank of libscgbm (scgbm_comm*)."""

    def __init__(self, handle, rank, nranks, device, allreduce_max_int=None):
        self.handle, self.rank, self.nranks, self.device = handle, rank, nranks, device
        self._armi = allreduce_max_int

    def allreduce_max_int(self, v: int) -> int:
        if self._armi is not None:
            return int(self._armi(int(v)))
        buf = np.array([float(v)])
        L.check(L.load().scgbm_comm_allreduce_f64(self.handle, L.ptr(buf), 1, 1))
        return int(buf[0])

    def allreduce(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        buf = np.ascontiguousarray(arr, dtype=np.float64).copy()
        L.check(L.load().scgbm_comm_allreduce_f64(self.handle, L.ptr(buf), buf.size, 1 if op == "max" else 0))
        return buf

    def allgather_bytes(self, payload: bytes) -> list:
        """Every rank's payload, in rank order, on every rank -- over the library's own NCCL all-reduce
        (lengths first, then each rank's bytes in its own slot of a zero vector), so no second transport is
        needed for small host-side agreements such as the batch levels."""
        lens = np.zeros(self.nranks)
        lens[self.rank] = len(payload)
        lens = self.allreduce(lens).astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        buf = np.zeros(max(int(offs[-1]), 1))
        buf[offs[self.rank]:offs[self.rank + 1]] = np.frombuffer(payload, dtype=np.uint8)
        buf = self.allreduce(buf).astype(np.uint8)
        return [buf[offs[r]:offs[r + 1]].tobytes() for r in range(self.nranks)]

    def destroy(self):
        if self.handle:
            L.load().scgbm_comm_destroy(self.handle)
            self.handle = None


def make_unique_id() -> bytes:
    buf = (C.c_uint8 * L.UNIQUE_ID_BYTES)()
    L.check(L.load().scgbm_comm_unique_id(buf))
    return bytes(buf)


def init_comm(rank: int, nranks: int, device: int, broadcast_bytes) -> Comm:
    """`broadcast_bytes(payload_or_None) -> bytes` must return rank 0's payload on
    every rank (e.g. a torch.distributed or MPI broadcast; tests/dist_util.py has the torch one)."""
    uid = broadcast_bytes(make_unique_id() if rank == 0 else None)
    if len(uid) != L.UNIQUE_ID_BYTES:
        raise ValueError("bad NCCL unique id")
    h = C.c_void_p()
    ub = (C.c_uint8 * L.UNIQUE_ID_BYTES).from_buffer_copy(uid)
    L.check(L.load().scgbm_comm_init_rank(C.byref(h), ub, nranks, rank, device))
    return Comm(h, rank, nranks, device)

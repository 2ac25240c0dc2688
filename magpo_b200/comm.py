"""Data-parallel communicator of the B200 path: one process per GPU, NCCL over NVLink / NVSwitch through the library's own
`magpo_comm_*` entry points (include/magpo_b200.h) — the `jax.lax.pmean(..., "device")` of rec_magpo.py:399-409. torch.distributed
is not involved in the exchange; the launcher's TCP key-value store is only the host channel that carries the 128-byte NCCL unique
id from rank 0 to the other ranks (the same store `torchrun` workers use for rendezvous).
"""
from __future__ import annotations

import ctypes as C
import os
from datetime import timedelta

import torch

from . import _lib as L


def _store(rank: int, world_size: int, addr: str, port: int, timeout_s: int = 600):
    """The rendezvous store of an env:// launch: hosted by the torchrun agent (workers are clients), else by rank 0."""
    from torch.distributed import TCPStore

    agent_store = os.environ.get("TORCHELASTIC_USE_AGENT_STORE", "False") == "True"
    return TCPStore(addr, port, world_size, is_master=(rank == 0 and not agent_store), timeout=timedelta(seconds=timeout_s),
                    wait_for_workers=False)


class NcclComm:
    """`MagpoComm` of this rank. `allreduce_sum / allreduce_max` are in place on float32 device tensors and enqueue on the current
    stream; `attach(learner)` lets `magpo_minibatch_grads` reduce the gradients itself, overlapped with the backward."""

    def __init__(self, rank: int, world_size: int, device, store=None, key: str = "magpo/nccl_unique_id"):
        self.rank, self.world_size, self.dev = rank, world_size, torch.device(device)
        lib = L.lib()
        if world_size > 1 and not lib.magpo_comm_available():
            raise L.MagpoError("libnccl.so.2 could not be loaded: no multi-GPU exchange (there is no fallback)")
        uid = C.create_string_buffer(128)
        if world_size > 1:
            if store is None:
                store = _store(rank, world_size, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))
            if rank == 0:
                L.check(lib.magpo_comm_unique_id(uid), "magpo_comm_unique_id")
                store.set(key, uid.raw)
            else:
                uid = C.create_string_buffer(bytes(store.get(key)), 128)
        self._handle = L.vp()
        with torch.cuda.device(self.dev):
            L.check(lib.magpo_comm_init(world_size, rank, uid, C.byref(self._handle)), "magpo_comm_init")
        self._one = torch.zeros(1, dtype=torch.float32, device=self.dev)

    @classmethod
    def from_env(cls, device=None) -> "NcclComm":
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        return cls(rank, world, device if device is not None else torch.device("cuda", local))

    @property
    def handle(self):
        return self._handle

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
        L.check(L.lib().magpo_comm_allreduce_sum(self._handle, L.stream_ptr(), L.ptr(t), C.c_int64(t.numel())), "magpo_comm_allreduce_sum")
        return t

    def allreduce_max(self, t: torch.Tensor) -> torch.Tensor:
        assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
        L.check(L.lib().magpo_comm_allreduce_max(self._handle, L.stream_ptr(), L.ptr(t), C.c_int64(t.numel())), "magpo_comm_allreduce_max")
        return t

    def barrier(self) -> None:
        """All ranks have reached this point and this device is idle."""
        if self.world_size > 1:
            self.allreduce_sum(self._one.zero_())
        torch.cuda.synchronize(self.dev)

    def attach(self, learner) -> None:
        L.check(L.lib().magpo_context_set_comm(learner.ctx, self._handle), "magpo_context_set_comm")
        learner.comm, learner.world_size = self, self.world_size

    def close(self) -> None:
        if self._handle:
            torch.cuda.synchronize(self.dev)
            L.lib().magpo_comm_destroy(self._handle)
            self._handle = L.vp()

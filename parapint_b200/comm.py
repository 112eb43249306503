"""Collectives of the Schur-complement solve, over ``torch.distributed``.

The reference issues MPI collectives on ``COMM_WORLD``
(``mpi_explicit_schur_complement.py:21,343,387,427-429``).  Here one process
drives one GPU and the same reductions run as NCCL all-reduces on device
buffers (NVLink 5 / NVSwitch); with the ``gloo`` backend the identical code path
runs on CPU tensors, which is how the multi-rank host logic is tested without
GPUs.  A process that never initialised ``torch.distributed`` is a 1-rank job.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

PEER_MAX_DOUBLES = 16384     # larger payloads (32 MB at config 5) are bandwidth-bound: NCCL


class PeerBuffers:
    """Two symmetric (peer-mapped) contribution buffers used alternately + a local result buffer: the operands of
    the one-shot peer all-reduce (``pp_peer_allreduce``, ``csrc/peer.cuh``)."""

    def __init__(self, comm, n, device, slot):
        import torch.distributed._symmetric_memory as symm_mem
        from . import native
        self.comm, self.n, self.slot = comm, int(n), int(slot)
        self.lib = native.load()
        group = comm.group if comm.group is not None else dist.group.WORLD
        self.bufs, self.tables = [], []
        for _ in range(2):
            t = symm_mem.empty(self.n, dtype=torch.float64, device=device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, group)
            if len(hdl.buffer_ptrs) != comm.size or hdl.signal_pad_size < 4 * (self.slot + comm.size):
                raise RuntimeError("symmetric memory: unexpected world size or signal pad size")
            ptrs = (C.c_void_p * comm.size)(*[int(p) for p in hdl.buffer_ptrs])
            pads = (C.c_void_p * comm.size)(*[int(p) for p in hdl.signal_pad_ptrs])
            self.bufs.append(t)
            self.tables.append((ptrs, pads, hdl))
        self.out = torch.zeros(self.n, dtype=torch.float64, device=device)
        self.turn = 0
        self.seq = 0

    def current(self):
        return self.bufs[self.turn]

    def reduce(self):
        """Sum of every rank's current buffer -> ``self.out``; the other buffer becomes current."""
        ptrs, pads, _ = self.tables[self.turn]
        self.seq += 1
        stream = torch.cuda.current_stream(self.out.device).cuda_stream
        code = self.lib.pp_peer_allreduce(self.comm.size, self.comm.rank, ptrs, pads, self.slot, self.seq & 0xffffffff,
                                          self.n, C.c_void_p(self.out.data_ptr()), C.c_void_p(stream))
        if code != 0:
            from . import native
            raise RuntimeError(f"pp_peer_allreduce failed: {native.last_error()}")
        self.turn ^= 1
        return self.out


class Communicator:
    def __init__(self, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.size = dist.get_world_size(group) if self.active else 1
        self._peer = {}          # data_ptr of a symmetric contribution buffer -> its PeerBuffers
        self._pool = {}          # size -> released PeerBuffers (every rank releases and re-acquires in the same order)
        self._slots = 0
        self.peer_enabled = (self.size > 1 and self.active and os.environ.get("PARAPINT_B200_PEER_ALLREDUCE", "1") != "0"
                             and dist.get_backend(group) == "nccl")
        self.peer_error = None

    def exchange_buffers(self, n, device):
        """Buffers for a SUM exchange of ``n`` doubles: a list of one ordinary device tensor (reduced in place by
        NCCL), or -- small payloads on NVLink-connected GPUs -- two symmetric tensors to be used alternately, whose
        reduction is the one-shot peer all-reduce.  Collective (every rank calls it with the same ``n``)."""
        if self.peer_enabled and 0 < n <= PEER_MAX_DOUBLES:
            try:
                pool = self._pool.setdefault(int(n), [])
                if pool:                      # released by an earlier solver / symbolic phase: same size, same channel
                    pb = pool.pop()
                else:
                    pb = PeerBuffers(self, n, device, self._slots)
                    self._slots += self.size
                    for t in pb.bufs:
                        self._peer[t.data_ptr()] = pb
                return pb.bufs
            except Exception as e:  # noqa: BLE001 - no peer access / symmetric memory unavailable: NCCL does it
                self.peer_error = repr(e)
                self.peer_enabled = False
        return [torch.zeros(max(int(n), 1), dtype=torch.float64, device=device)]

    def release_buffers(self, bufs):
        """Give peer-mapped exchange buffers back (a solver that re-analyses or goes away); collective in effect:
        every rank must release and acquire in the same order."""
        if bufs and self._peer:
            pb = self._peer.get(bufs[0].data_ptr())
            if pb is not None and pb not in self._pool.setdefault(pb.n, []):
                self._pool[pb.n].append(pb)

    def allreduce_sum_(self, tensor: torch.Tensor) -> torch.Tensor:
        """SUM all-reduce (S values ``:343``; coupling rhs ``:387``; inertia ``:427-429``).  Returns the tensor that
        holds the sum: ``tensor`` itself (NCCL / gloo, in place), or the local result buffer of the peer exchange
        when ``tensor`` is one of the symmetric buffers handed out by :meth:`exchange_buffers`."""
        if self.size > 1:
            pb = self._peer.get(tensor.data_ptr()) if self._peer else None
            if pb is not None and pb.current() is tensor:
                return pb.reduce()
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_max_(self, tensor: torch.Tensor) -> torch.Tensor:
        """In-place MAX all-reduce (status agreement, role of ``_gather_results`` ``:19-30``)."""
        if self.size > 1:
            dist.all_reduce(tensor, op=dist.ReduceOp.MAX, group=self.group)
        return tensor

    def allgather_object(self, obj):
        """Every rank's ``obj`` in rank order (border structure of all blocks for the pattern of a sparse Schur
        complement, ``:244-247``; symbolic phase only)."""
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)

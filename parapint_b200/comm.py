"""Collectives of the Schur-complement solve, over ``torch.distributed``.

The reference issues MPI collectives on ``COMM_WORLD``
(``mpi_explicit_schur_complement.py:21,343,387,427-429``).  Here one process
drives one GPU and the same reductions run as NCCL all-reduces on device
buffers (NVLink 5 / NVSwitch); with the ``gloo`` backend the identical code path
runs on CPU tensors, which is how the multi-rank host logic is tested without
GPUs.  A process that never initialised ``torch.distributed`` is a 1-rank job.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class Communicator:
    def __init__(self, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.size = dist.get_world_size(group) if self.active else 1

    def allreduce_sum_(self, tensor: torch.Tensor) -> torch.Tensor:
        """In-place SUM all-reduce (S values ``:343``; coupling rhs ``:387``; inertia ``:427-429``)."""
        if self.size > 1:
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

"""The reference CPU path on all host cores: a process-per-rank emulation of
``MPISchurComplementLinearSolver`` (TEST / BENCH INFRASTRUCTURE, see ``oracle/__init__.py``).

``mpirun`` and mpi4py do not exist in this image, so the partition of
``mpi_explicit_schur_complement.py:198-203`` (round-robin blocks over ranks,
``mpi_sc_ip_interface.py:14-19``) is run with ``multiprocessing``: every worker
process owns the blocks ``i % P == rank`` and executes the oracle restatement of
the reference arithmetic on them (SciPy SuperLU leaves, one back-solve per
nonzero border row); the parent plays the Allreduce (``:343,387``) and the
replicated coupling factorisation (``:352-360``).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from oracle.schur_oracle import SUCCESSFUL, WARNING, SchurOracle, _worst


def _worker(conn, kkt, rhs, blocks):
    os.environ["OMP_NUM_THREADS"] = "1"
    o = SchurOracle()
    o.symbolic(kkt, blocks)
    while True:
        cmd, payload = conn.recv()
        if cmd == "numeric":
            conn.send(o.local_contribution(kkt))
        elif cmd == "forward":
            conn.send(o.local_forward(rhs))
        elif cmd == "backward":
            from parapint_b200.carriers import BlockVector
            out = BlockVector(o.nblocks + 1)
            o.local_backward(rhs, payload, out)
            conn.send({i: np.asarray(out.get_block(i)) for i in blocks})
        elif cmd == "stop":
            conn.close()
            return


class PartitionedCpuSolver:
    """Factor + solve with ``procs`` worker processes (fork; inputs are inherited, not pickled)."""

    def __init__(self, kkt, rhs, procs=None):
        self.kkt, self.rhs = kkt, rhs
        self.N = kkt.bshape[0] - 1
        self.procs = max(1, min(procs or os.cpu_count() or 1, self.N))
        ctx = mp.get_context("fork")
        self.workers = []
        for r in range(self.procs):
            parent, child = ctx.Pipe()
            blocks = [i for i in range(self.N) if i % self.procs == r]
            p = ctx.Process(target=_worker, args=(child, kkt, rhs, blocks), daemon=True)
            p.start()
            child.close()
            self.workers.append((p, parent))
        self.coupling = SchurOracle()
        self.coupling.symbolic(kkt, [])

    def _all(self, cmd, payload=None):
        for _, c in self.workers:
            c.send((cmd, payload))
        return [c.recv() for _, c in self.workers]

    def factor_and_solve(self):
        """One 'KKT factor + solve': numeric factorisation then one back-solve. Returns (status, x blocks, x_c)."""
        parts = self._all("numeric")
        status = SUCCESSFUL
        total = None
        for st, contrib in parts:
            status = _worst(status, st)
            total = contrib if total is None else total + contrib
        if status not in (SUCCESSFUL, WARNING):
            return status, None, None
        self.coupling.status = SUCCESSFUL
        status = self.coupling.factor_coupling(self.kkt, total)
        r_c = sum(self._all("forward"))
        x_c = self.coupling.solve_coupling(self.rhs, r_c)
        blocks = {}
        for part in self._all("backward", x_c):
            blocks.update(part)
        return status, blocks, x_c

    def close(self):
        for p, c in self.workers:
            try:
                c.send(("stop", None))
            except Exception:  # noqa: BLE001
                pass
        for p, _ in self.workers:
            p.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def time_serial(kkt, rhs, min_seconds=8.0, max_reps=20):
    """The serial reference algorithm (1 core): median seconds of numeric + solve over repeated runs."""
    o = SchurOracle()
    o.symbolic(kkt)
    times = []
    t_begin = time.perf_counter()
    x = None
    while len(times) < max_reps and (time.perf_counter() - t_begin < min_seconds or len(times) < 2):
        t0 = time.perf_counter()
        assert o.numeric(kkt) == SUCCESSFUL
        x = o.solve(rhs)
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), len(times), x

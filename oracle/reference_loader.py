"""Load the UNMODIFIED reference hot-path files from ``/root/reference``.

TEST INFRASTRUCTURE, build-container only (``/root/reference`` does not exist on
the GPU box).  ``import parapint`` fails here because Pyomo / mpi4py are not
installed, but the files that make up the hot path only touch a
small surface of those packages.  This module registers
stand-ins for that surface in ``sys.modules`` and then executes the reference
files *in place, by path* -- nothing is copied into this repository.

Used by ``tests/golden/make_golden.py`` (fixture generation) and by the
``reference``-marked CPU tests that pin ``oracle.schur_oracle`` and
``oracle.kkt_generator`` against the reference itself when it is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from parapint_b200 import carriers

REFERENCE_ROOT = os.environ.get("PARAPINT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "parapint", "linalg", "results.py"))


class _Timer:
    def start(self, name):
        pass

    def stop(self, name):
        pass


class _Comm:
    """One-rank stand-in for ``mpi4py.MPI.COMM_WORLD``."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def Split(self, color, key):
        return self

    def allgather(self, obj):
        return [obj]

    def allreduce(self, obj, op=None):
        return obj

    def Allreduce(self, send, recv, op=None):
        np.copyto(recv, send)

    def Allgatherv(self, send, recv):
        buf = recv[0] if isinstance(recv, (list, tuple)) else recv
        np.copyto(buf, send)


class _MPIBlockMatrix(carriers.BlockMatrix):
    def __init__(self, nbrows, nbcols, rank_ownership, mpi_comm, assert_correct_owners=False):
        super().__init__(nbrows, nbcols)
        self.rank_ownership = np.asarray(rank_ownership, dtype=np.int64)
        self.mpi_comm = mpi_comm

    def broadcast_block_sizes(self):
        pass

    def to_local_array(self):
        return self.toarray()


class _MPIBlockVector(carriers.BlockVector):
    def __init__(self, nblocks, rank_owner, mpi_comm, assert_correct_owners=False):
        super().__init__(nblocks)
        self.rank_ownership = np.asarray(rank_owner, dtype=np.int64)
        self.mpi_comm = mpi_comm

    def broadcast_block_sizes(self):
        pass

    def make_local_copy(self):
        return self.copy()

    def copy_structure(self):
        out = _MPIBlockVector(self.nblocks, self.rank_ownership, self.mpi_comm)
        base = carriers.BlockVector.copy_structure(self)
        for i in range(self.nblocks):
            if base.get_block(i) is not None:
                out.set_block(i, base.get_block(i))
        return out


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stand_ins():
    if "pyomo" in sys.modules and not getattr(sys.modules["pyomo"], "_parapint_b200_stub", False):
        return  # a real pyomo is present; use it
    _module("pyomo", _parapint_b200_stub=True)
    _module("pyomo.common")
    _module("pyomo.common.timing", HierarchicalTimer=_Timer)
    _module("pyomo.contrib")
    _module("pyomo.contrib.pynumero")
    _module("pyomo.contrib.pynumero.sparse", BlockMatrix=carriers.BlockMatrix, BlockVector=carriers.BlockVector)
    _module("pyomo.contrib.pynumero.sparse.block_matrix", BlockMatrix=carriers.BlockMatrix)
    _module("pyomo.contrib.pynumero.sparse.block_vector", BlockVector=carriers.BlockVector)
    _module("pyomo.contrib.pynumero.sparse.mpi_block_matrix", MPIBlockMatrix=_MPIBlockMatrix)
    _module("pyomo.contrib.pynumero.sparse.mpi_block_vector", MPIBlockVector=_MPIBlockVector)
    mpi = _module("mpi4py.MPI", COMM_WORLD=_Comm(), Comm=_Comm, MAX="max", SUM="sum")
    _module("mpi4py", MPI=mpi)


def _exec(dotted, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = None


def load():
    """Return a namespace with the reference's own classes (executed from /root/reference)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stand_ins()
    for pkg in ("parapint", "parapint.linalg", "parapint.linalg.schur_complement", "parapint.examples",
                "parapint.examples.performance", "parapint.examples.performance.schur_complement"):
        mod = types.ModuleType(pkg)
        mod.__path__ = []
        sys.modules[pkg] = mod
    results = _exec("parapint.linalg.results", "parapint/linalg/results.py")
    base = _exec("parapint.linalg.base_linear_solver_interface", "parapint/linalg/base_linear_solver_interface.py")
    scipy_leaf = _exec("parapint.linalg.scipy_interface", "parapint/linalg/scipy_interface.py")
    serial = _exec("parapint.linalg.schur_complement.explicit_schur_complement",
                   "parapint/linalg/schur_complement/explicit_schur_complement.py")
    mpi = _exec("parapint.linalg.schur_complement.mpi_explicit_schur_complement",
                "parapint/linalg/schur_complement/mpi_explicit_schur_complement.py")
    linalg = sys.modules["parapint.linalg"]
    linalg.LinearSolverInterface = base.LinearSolverInterface
    linalg.LinearSolverStatus = results.LinearSolverStatus
    linalg.LinearSolverResults = results.LinearSolverResults
    linalg.ScipyInterface = scipy_leaf.ScipyInterface
    linalg.SchurComplementLinearSolver = serial.SchurComplementLinearSolver
    linalg.MPISchurComplementLinearSolver = mpi.MPISchurComplementLinearSolver
    linalg.InteriorPointMA27Interface = None  # HSL MA27 binary absent from this image
    linalg.MumpsInterface = None  # pymumps absent from this image
    sys.modules["parapint"].linalg = linalg
    _exec("parapint.examples.performance.schur_complement.utils",
          "parapint/examples/performance/schur_complement/utils.py")
    gen = _exec("parapint.examples.performance.schur_complement.create_model",
                "parapint/examples/performance/schur_complement/create_model.py")
    driver = _exec("parapint.examples.performance.schur_complement.main",
                   "parapint/examples/performance/schur_complement/main.py")
    ns = types.SimpleNamespace(
        LinearSolverStatus=results.LinearSolverStatus,
        LinearSolverResults=results.LinearSolverResults,
        LinearSolverInterface=base.LinearSolverInterface,
        ScipyInterface=scipy_leaf.ScipyInterface,
        SchurComplementLinearSolver=serial.SchurComplementLinearSolver,
        MPISchurComplementLinearSolver=mpi.MPISchurComplementLinearSolver,
        Model=gen.Model,
        MPIModel=gen.MPIModel,
        perf_main=driver,
        BlockMatrix=carriers.BlockMatrix,
        BlockVector=carriers.BlockVector,
        MPIBlockMatrix=_MPIBlockMatrix,
        MPIBlockVector=_MPIBlockVector,
        comm=sys.modules["mpi4py.MPI"].COMM_WORLD,
    )
    _cache = ns
    return ns

"""B200-native Schur-complement KKT linear solver behind parapint's ``LinearSolverInterface``.

Public surface (mirrors what ``parapint.linalg`` exports for this path):

* :class:`B200SchurComplementLinearSolver` -- replaces ``SchurComplementLinearSolver`` /
  ``MPISchurComplementLinearSolver`` and their sub-solvers.
* :class:`LinearSolverInterface`, :class:`LinearSolverResults`, :class:`LinearSolverStatus`.
* :class:`BlockMatrix`, :class:`BlockVector` -- PyNumero-compatible carriers for use without Pyomo.
"""
from .carriers import BlockMatrix, BlockVector
from .interface import LinearSolverInterface, LinearSolverResults, LinearSolverStatus
from .comm import Communicator
from .schur_solver import B200SchurComplementLinearSolver, CudaBackend

__all__ = [
    "B200SchurComplementLinearSolver", "CudaBackend", "Communicator", "LinearSolverInterface",
    "LinearSolverResults", "LinearSolverStatus", "BlockMatrix", "BlockVector",
]

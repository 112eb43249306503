"""The linear-solver plugin contract this package drops in behind.

Mirrors reference ``parapint/linalg/base_linear_solver_interface.py:5-56``
(``LinearSolverInterface``) and ``parapint/linalg/results.py:4-14``
(``LinearSolverStatus`` / ``LinearSolverResults``): same names, same argument
meaning, same status codes.  When the real ``parapint`` package is importable
its classes are re-exported, so that ``isinstance`` checks inside
``parapint.algorithms.interior_point`` hold for the solver defined here; when
it is not (Pyomo is absent from this image) the equivalents below are used.
"""
from __future__ import annotations

import enum
import logging
from abc import ABC, abstractmethod

try:  # pragma: no cover - parapint is not installed in the build image
    from parapint.linalg.base_linear_solver_interface import LinearSolverInterface  # type: ignore
    from parapint.linalg.results import LinearSolverResults, LinearSolverStatus  # type: ignore
except Exception:  # noqa: BLE001

    class LinearSolverStatus(enum.Enum):
        successful = 0
        not_enough_memory = 1
        singular = 2
        error = 3
        warning = 4

    class LinearSolverResults:
        def __init__(self, status=None):
            self.status = status

    class LinearSolverInterface(ABC):
        """Factor / inertia / solve contract used by ``ip_solve``
        (callers: reference ``parapint/algorithms/interior_point.py:347,371,387,544,566,646``)."""

        @classmethod
        def getLoggerName(cls):
            return "linear_solver"

        @classmethod
        def getLogger(cls):
            return logging.getLogger("algorithms." + cls.getLoggerName())

        @abstractmethod
        def do_symbolic_factorization(self, matrix, raise_on_error=True, timer=None):
            """Analyse the nonzero structure of ``matrix``."""

        @abstractmethod
        def do_numeric_factorization(self, matrix, raise_on_error=True, timer=None):
            """Factorize ``matrix``; only valid after the symbolic phase."""

        def increase_memory_allocation(self, factor):
            raise NotImplementedError("Should be implemented by base class.")

        @abstractmethod
        def do_back_solve(self, rhs):
            """Solve ``matrix * x = rhs``; only valid after the numeric phase."""

        @abstractmethod
        def get_inertia(self):
            """(num_pos, num_neg, num_zero) of the factorized matrix."""


__all__ = ["LinearSolverInterface", "LinearSolverResults", "LinearSolverStatus"]

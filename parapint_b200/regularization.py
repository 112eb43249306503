"""Device-side inertia correction: the regularised KKT matrix as *base matrix + three diagonal shifts*.

parapint's inertia-correction loop (reference ``parapint/algorithms/interior_point.py:369-395``) asks the problem
interface for a regularised copy of the KKT matrix before every retry::

    kkt = interface.regularize_equality_gradient(kkt=kkt, coef=-delta, copy_kkt=False)   # interface.py:590-608
    kkt = interface.regularize_hessian(kkt=kkt, coef=delta, copy_kkt=False)              # interface.py:610-619
    linear_solver.do_numeric_factorization(matrix=kkt, ...)

and a solver that only sees the resulting ``BlockMatrix`` has to gather and upload all of its values again -- and to
repeat its symbolic phase, because ``hess += delta * I`` changes the COO pattern.  The shifts themselves are three
numbers.  :class:`DeviceRegularizationMixin` overrides the two ``regularize_*`` methods of an interface so that they
return a :class:`RegularizedKKT` -- the unmodified matrix plus the shift values -- and
``B200SchurComplementLinearSolver`` then adds the shifts on the device (``pp_set_shifts``) and factorises the values
it already holds (``PP_VALUES_REUSE``): a retry costs a refactorisation only.  Nothing changes in ``ip_solve``.

Semantics reproduced (they differ per class, and the first one accumulates over the retries of one iteration):

=====  ==========================================================  ==========================================
class  rows                                                        value after ``regularize_*``
=====  ==========================================================  ==========================================
1      primal variables of every block (Hessian diagonal)          base + sum of the ``coef`` of every
                                                                   ``regularize_hessian`` call (``hess += ptb``)
2      equality / inequality / linking multipliers (also the       ``coef`` of the last
       coupling-side forward multipliers of the dynamic layout)    ``regularize_equality_gradient`` (blocks
                                                                   that hold ``0 * I`` are *replaced*)
3      coupling variables (first-stage variables, interface        ``coef`` of the last ``regularize_hessian``
       states)                                                     (``kkt.set_block(N, N, coef * I)``)
=====  ==========================================================  ==========================================
"""
from __future__ import annotations

import numpy as np

HESSIAN, MULTIPLIER, COUPLING_PRIMAL = 1, 2, 3


class RegularizedKKT:
    """``base`` (never modified) + diagonal shifts.  Quacks like the base matrix for everything else."""

    def __init__(self, base, classes, shifts=(0.0, 0.0, 0.0), token=None):
        self.base = base
        self.classes = classes            # (list of int8 arrays per diagonal block index, int8 array for the coupling rows)
        self.shifts = tuple(float(v) for v in shifts)
        self.token = token                # identifies the evaluation the values of `base` belong to

    def copy(self, deep=True):
        return RegularizedKKT(self.base, self.classes, self.shifts, self.token)

    def with_hessian(self, coef):
        return RegularizedKKT(self.base, self.classes, (self.shifts[0] + coef, self.shifts[1], coef), self.token)

    def with_multipliers(self, coef):
        return RegularizedKKT(self.base, self.classes, (self.shifts[0], coef, self.shifts[2]), self.token)

    def __getattr__(self, name):          # bshape, get_block, shape, ...: the structure is the base matrix's
        return getattr(self.base, name)

    def materialize(self):
        """The regularised matrix as an ordinary block matrix (for a solver that knows nothing of this class)."""
        import scipy.sparse as sp
        out = self.base.copy()
        N = out.bshape[0] - 1
        per_block, coupling = self.classes
        sh = np.array((0.0,) + self.shifts)
        for i in range(N + 1):
            cls = coupling if i == N else per_block.get(i) if isinstance(per_block, dict) else per_block[i]
            blk = out.get_block(i, i)
            if cls is None or blk is None:
                continue
            d = sh[np.asarray(cls, dtype=np.int64)]
            n = d.size
            cur = blk.tocoo()
            out.set_block(i, i, sp.coo_matrix((np.concatenate([cur.data, d]),
                                               (np.concatenate([cur.row, np.arange(n)]),
                                                np.concatenate([cur.col, np.arange(n)]))), shape=cur.shape))
        return out


class DeviceRegularizationMixin:
    """Mix into a parapint-style interface (before the interface class in the MRO)::

        class MyInterface(DeviceRegularizationMixin, StochasticSchurComplementInteriorPointInterface): ...

    The interface must provide ``regularization_classes()`` -> ``(per_block, coupling)``: for every diagonal block
    index an ``int8`` array with the class (0 none, 1, 2, 3 -- see the module docstring) of each of its rows, and one
    array for the coupling rows.  ``evaluate_primal_dual_kkt_matrix`` is wrapped to tag every evaluation with a token;
    ``regularize_*`` must be handed the matrix of the latest evaluation (or a copy / a ``RegularizedKKT`` of it), which
    is how ``ip_solve`` uses them (``interior_point.py:383-387``)."""

    _pp_token = 0

    def evaluate_primal_dual_kkt_matrix(self, *args, **kwargs):
        kkt = super().evaluate_primal_dual_kkt_matrix(*args, **kwargs)
        self._pp_token += 1
        try:
            kkt._pp_token = (id(self), self._pp_token)
        except AttributeError:  # a matrix class with __slots__: the solver then falls back to object identity
            pass
        return kkt

    def _pp_wrap(self, kkt):
        if isinstance(kkt, RegularizedKKT):
            return kkt
        return RegularizedKKT(kkt, self.regularization_classes(), token=(id(self), self._pp_token))

    def regularize_equality_gradient(self, kkt, coef, copy_kkt=True):
        return self._pp_wrap(kkt).with_multipliers(coef)

    def regularize_hessian(self, kkt, coef, copy_kkt=True):
        return self._pp_wrap(kkt).with_hessian(coef)

"""Flat restatement of parapint's interior-point driver and of the pieces of its problem interfaces
that build the block-bordered KKT system (TEST INFRASTRUCTURE, see ``oracle/__init__.py``).

``import parapint`` is impossible in this image (no Pyomo / ASL), so the parts of the reference that
*call* the linear-solver plugin are restated here with closed-form models, in order to check the plugin
end to end (SURVEY.md 8(d) "end-to-end inputs"): the same IPM run with the reference algorithm's
SciPy-leaf Schur solver and with the B200 solver must take the same number of iterations and reach the
same objective.  Followed line by line:

* ``ip_solve`` -- ``parapint/algorithms/interior_point.py:405-631``; ``check_convergence`` ``:174-317``;
  inertia correction ``numeric_factorization`` ``:337-402``; ``try_factorization_and_reallocation``
  ``:634-652``; ``fraction_to_the_boundary`` ``:655-758``; ``process_init*`` ``:761-799``; defaults of
  ``IPOptions`` ``:159-171`` / ``InertiaCorrectionOptions`` ``:57-60``.
* single-NLP KKT / rhs / dual steps / regularisation -- ``parapint/interfaces/interface.py:432-528,
  536-570, 590-619`` (order ``[x, s, lam_eq, lam_ineq]``, duplicated barrier diagonal appended to the
  Hessian COO, explicit ``0*I`` blocks).
* two-stage stochastic structure -- ``parapint/interfaces/schur_complement/sc_ip_interface.py:1245-1317``
  (nested diagonal blocks ``[[KKT_i, L_i^T],[L_i, 0*I]]``, border ``[0 | -C_i^T]``, ``Q = 0*I``), rhs
  ``:1677-1693``, solution unpacking ``:1695-1710``, regularisation ``:1736-1757``.

Vectors are flat numpy arrays here (the reference uses PyNumero BlockVectors); the KKT matrix and its
right-hand side keep the reference's nested block structure because that is the plugin's input format.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from parapint_b200.carriers import BlockMatrix, BlockVector
from parapint_b200.interface import LinearSolverStatus


class QuadraticScenario:
    """min 0.5 x'Hx + c'x  s.t.  A_eq x = b_eq,  g_lb <= A_in x <= g_ub,  lb <= x <= ub  (closed-form NLP:
    the role PyomoNLP plays at ``interface.py:251-256``).  Primals start at zero, as an uninitialised Pyomo
    model does; duals start at zero."""

    def __init__(self, c, lb, ub, A_in, g_lb, g_ub, A_eq=None, b_eq=None, H=None):
        self.c = np.asarray(c, dtype=float)
        self.n = self.c.size
        self.lb = np.asarray(lb, dtype=float)
        self.ub = np.asarray(ub, dtype=float)
        self.A_in = sp.coo_matrix(A_in, dtype=float).reshape((-1, self.n)) if A_in is not None else sp.coo_matrix((0, self.n))
        self.g_lb = np.asarray(g_lb, dtype=float)
        self.g_ub = np.asarray(g_ub, dtype=float)
        self.A_eq = sp.coo_matrix(A_eq, dtype=float) if A_eq is not None else sp.coo_matrix((0, self.n))
        self.b_eq = np.asarray(b_eq, dtype=float) if b_eq is not None else np.zeros(0)
        self.H = sp.coo_matrix(H, dtype=float) if H is not None else sp.coo_matrix((self.n, self.n))
        self.n_eq, self.n_in = self.A_eq.shape[0], self.A_in.shape[0]
        self.x = np.zeros(self.n)

    def objective(self):
        return float(self.c @ self.x + 0.5 * self.x @ (self.H @ self.x))

    def grad(self):
        return self.c + self.H @ self.x

    def eq(self):
        return self.A_eq @ self.x - self.b_eq

    def ineq(self):
        return self.A_in @ self.x


class ScenarioInterface:
    """Restates ``InteriorPointInterface`` (``interface.py:250-619``) for one closed-form scenario."""

    def __init__(self, nlp: QuadraticScenario):
        self.nlp = nlp
        self.relax = 0.0
        self.slacks = nlp.ineq().copy()                      # :261, :310-312
        self.duals_eq = np.zeros(nlp.n_eq)
        self.duals_ineq = np.zeros(nlp.n_in)
        self.init_zl = np.ones(nlp.n)                        # :626-631 (no ipopt suffixes)
        self.init_zu = np.ones(nlp.n)
        self.init_zl[np.isneginf(nlp.lb)] = 0                # :267-268
        self.init_zu[np.isinf(nlp.ub)] = 0
        self.zl, self.zu = self.init_zl.copy(), self.init_zu.copy()
        self.init_sl = np.zeros(nlp.n_in)                    # :276-281 with zero initial multipliers
        self.init_su = np.zeros(nlp.n_in)
        self.sl, self.su = self.init_sl.copy(), self.init_su.copy()
        self.barrier = None
        self.d_x = self.d_s = self.d_eq = self.d_in = None

    # bounds with relaxation, :395-425
    def _relaxed(self, b, sign):
        if self.relax == 0:
            return b
        return b + sign * self.relax * np.maximum(1.0, np.abs(b))

    def x_lb(self): return self._relaxed(self.nlp.lb, -1.0)
    def x_ub(self): return self._relaxed(self.nlp.ub, +1.0)
    def g_lb(self): return self._relaxed(self.nlp.g_lb, -1.0)
    def g_ub(self): return self._relaxed(self.nlp.g_ub, +1.0)

    def kkt(self):
        """``evaluate_primal_dual_kkt_matrix``, :432-491."""
        nlp = self.nlp
        x = nlp.x
        hess = nlp.H.tocoo()
        diag = self.zl / (x - self.x_lb()) + self.zu / (self.x_ub() - x)
        idx = np.arange(nlp.n)
        hess = sp.coo_matrix((np.concatenate([hess.data, diag]),
                              (np.concatenate([hess.row, idx]), np.concatenate([hess.col, idx]))), shape=(nlp.n, nlp.n))
        sdiag = self.sl / (self.slacks - self.g_lb()) + self.su / (self.g_ub() - self.slacks)
        ii = np.arange(nlp.n_in)
        slack_block = sp.coo_matrix((sdiag, (ii, ii)), shape=(nlp.n_in, nlp.n_in))
        eq_reg = sp.identity(nlp.n_eq, format="coo"); eq_reg.data.fill(0)
        in_reg = sp.identity(nlp.n_in, format="coo"); in_reg.data.fill(0)
        kkt = BlockMatrix(4, 4)
        kkt.set_block(0, 0, hess)
        kkt.set_block(1, 1, slack_block)
        kkt.set_block(2, 0, nlp.A_eq.tocoo())
        kkt.set_block(0, 2, nlp.A_eq.transpose().tocoo())
        kkt.set_block(3, 0, nlp.A_in.tocoo())
        kkt.set_block(0, 3, nlp.A_in.transpose().tocoo())
        kkt.set_block(3, 1, -sp.identity(nlp.n_in, format="coo"))
        kkt.set_block(1, 3, -sp.identity(nlp.n_in, format="coo"))
        kkt.set_block(2, 2, eq_reg)
        kkt.set_block(3, 3, in_reg)
        for k, sz in enumerate((nlp.n, nlp.n_in, nlp.n_eq, nlp.n_in)):
            kkt.set_row_size(k, sz)
            kkt.set_col_size(k, sz)
        return kkt

    def rhs(self):
        """``evaluate_primal_dual_kkt_rhs``, :493-528 (already negated)."""
        nlp = self.nlp
        x = nlp.x
        g1 = (nlp.grad() + nlp.A_eq.T @ self.duals_eq + nlp.A_in.T @ self.duals_ineq
              - self.barrier / (x - self.x_lb()) + self.barrier / (self.x_ub() - x))
        g2 = (-self.duals_ineq - self.barrier / (self.slacks - self.g_lb()) + self.barrier / (self.g_ub() - self.slacks))
        return [-g1, -g2, -nlp.eq(), -(nlp.ineq() - self.slacks)]

    def set_solution(self, blocks):
        self.d_x, self.d_s, self.d_eq, self.d_in = [np.asarray(b, dtype=float) for b in blocks]

    # dual steps, :548-570
    def d_zl(self): return (self.barrier - self.zl * self.d_x) / (self.nlp.x - self.x_lb()) - self.zl
    def d_zu(self): return (self.barrier + self.zu * self.d_x) / (self.x_ub() - self.nlp.x) - self.zu
    def d_sl(self): return (self.barrier - self.sl * self.d_s) / (self.slacks - self.g_lb()) - self.sl
    def d_su(self): return (self.barrier + self.su * self.d_s) / (self.g_ub() - self.slacks) - self.su


class StochasticInterface:
    """Restates ``StochasticSchurComplementInteriorPointInterface`` (``sc_ip_interface.py:1028-1757``) on flat
    vectors.  ``first_stage[i]`` lists the indices of scenario i's first-stage variables, in the order of the
    coupling variables."""

    def __init__(self, scenarios, first_stage):
        self.sc = [ScenarioInterface(s) for s in scenarios]
        self.N = len(self.sc)
        self.n_c = len(first_stage[0])
        self.L = [sp.coo_matrix((np.ones(len(fs)), (np.arange(len(fs)), np.asarray(fs))), shape=(len(fs), s.n))
                  for fs, s in zip(first_stage, scenarios)]                       # :1287-1301
        self.C = [sp.identity(self.n_c, format="coo") for _ in scenarios]        # :1303-1317 (all scenarios carry all)
        self.z = np.zeros(self.n_c)
        self.link_duals = [np.zeros(self.n_c) for _ in scenarios]
        self.d_link = [np.zeros(self.n_c) for _ in scenarios]
        self.d_z = np.zeros(self.n_c)
        nx = [s.nlp.n for s in self.sc]
        self.x_off = np.concatenate(([0], np.cumsum(nx)))
        self.s_off = np.concatenate(([0], np.cumsum([s.nlp.n_in for s in self.sc])))
        self.e_off = np.concatenate(([0], np.cumsum([s.nlp.n_eq + self.n_c for s in self.sc])))
        self.kkt_evals = 0

    # ---- sizes ----
    def n_eq_constraints(self): return int(self.e_off[-1])                       # :1437 (linking rows included)
    def n_ineq_constraints(self): return int(self.s_off[-1])
    def get_obj_factor(self): return 1.0

    def set_bounds_relaxation_factor(self, v):
        for s in self.sc:
            s.relax = v

    # ---- flat vectors ----
    def _cat_x(self, per_scenario, tail):
        return np.concatenate([*per_scenario, tail])

    def primals_lb(self): return self._cat_x([s.x_lb() for s in self.sc], np.full(self.n_c, -np.inf))
    def primals_ub(self): return self._cat_x([s.x_ub() for s in self.sc], np.full(self.n_c, np.inf))
    def ineq_lb(self): return np.concatenate([s.g_lb() for s in self.sc])
    def ineq_ub(self): return np.concatenate([s.g_ub() for s in self.sc])
    def init_primals(self): return self._cat_x([np.zeros(s.nlp.n) for s in self.sc], np.zeros(self.n_c))
    def init_slacks(self): return np.concatenate([s.nlp.A_in @ np.zeros(s.nlp.n) for s in self.sc])
    def init_duals_eq(self): return np.zeros(self.n_eq_constraints())
    def init_duals_ineq(self): return np.zeros(self.n_ineq_constraints())
    def init_duals_primals_lb(self): return self._cat_x([s.init_zl for s in self.sc], np.zeros(self.n_c))
    def init_duals_primals_ub(self): return self._cat_x([s.init_zu for s in self.sc], np.zeros(self.n_c))
    def init_duals_slacks_lb(self): return np.concatenate([s.init_sl for s in self.sc])
    def init_duals_slacks_ub(self): return np.concatenate([s.init_su for s in self.sc])

    def set_primals(self, v):
        for i, s in enumerate(self.sc):
            s.nlp.x = v[self.x_off[i]:self.x_off[i + 1]].copy()
        self.z = v[self.x_off[-1]:].copy()

    def set_slacks(self, v):
        for i, s in enumerate(self.sc):
            s.slacks = v[self.s_off[i]:self.s_off[i + 1]].copy()

    def set_duals_eq(self, v):
        for i, s in enumerate(self.sc):
            seg = v[self.e_off[i]:self.e_off[i + 1]]
            s.duals_eq = seg[: s.nlp.n_eq].copy()
            self.link_duals[i] = seg[s.nlp.n_eq:].copy()

    def set_duals_ineq(self, v):
        for i, s in enumerate(self.sc):
            s.duals_ineq = v[self.s_off[i]:self.s_off[i + 1]].copy()

    def _set_x_like(self, v, name):
        for i, s in enumerate(self.sc):
            setattr(s, name, v[self.x_off[i]:self.x_off[i + 1]].copy())

    def _set_s_like(self, v, name):
        for i, s in enumerate(self.sc):
            setattr(s, name, v[self.s_off[i]:self.s_off[i + 1]].copy())

    def set_duals_primals_lb(self, v): self._set_x_like(v, "zl")
    def set_duals_primals_ub(self, v): self._set_x_like(v, "zu")
    def set_duals_slacks_lb(self, v): self._set_s_like(v, "sl")
    def set_duals_slacks_ub(self, v): self._set_s_like(v, "su")

    def get_primals(self): return self._cat_x([s.nlp.x for s in self.sc], self.z)
    def get_slacks(self): return np.concatenate([s.slacks for s in self.sc])
    def get_duals_eq(self): return np.concatenate([np.concatenate([s.duals_eq, l]) for s, l in zip(self.sc, self.link_duals)])
    def get_duals_ineq(self): return np.concatenate([s.duals_ineq for s in self.sc])
    def get_duals_primals_lb(self): return self._cat_x([s.zl for s in self.sc], np.zeros(self.n_c))
    def get_duals_primals_ub(self): return self._cat_x([s.zu for s in self.sc], np.zeros(self.n_c))
    def get_duals_slacks_lb(self): return np.concatenate([s.sl for s in self.sc])
    def get_duals_slacks_ub(self): return np.concatenate([s.su for s in self.sc])

    # ---- evaluations ----
    def evaluate_objective(self): return sum(s.nlp.objective() for s in self.sc)
    def evaluate_grad_objective(self): return self._cat_x([s.nlp.grad() for s in self.sc], np.zeros(self.n_c))

    def evaluate_eq_constraints(self):
        return np.concatenate([np.concatenate([s.nlp.eq(), L @ s.nlp.x - C @ self.z])
                               for s, L, C in zip(self.sc, self.L, self.C)])

    def evaluate_ineq_constraints(self): return np.concatenate([s.nlp.ineq() for s in self.sc])

    # The scenarios are closed-form QPs: every constraint is linear, so both Jacobians are constant.  They are
    # assembled once (N x (N + 1) blocks through ``bmat`` cost more than the rest of an iteration) and the same CSR
    # object is handed out afterwards; callers only read it.
    def evaluate_jacobian_eq(self):
        jac = self.__dict__.get("_jac_eq")
        if jac is None:
            jac = self._jac_eq = self._assemble_jacobian_eq()
        return jac

    def evaluate_jacobian_ineq(self):
        jac = self.__dict__.get("_jac_ineq")
        if jac is None:
            jac = self._jac_ineq = self._assemble_jacobian_ineq()
        return jac

    def _assemble_jacobian_eq(self):
        rows = []
        for i, (s, L, C) in enumerate(zip(self.sc, self.L, self.C)):
            row = [None] * (self.N + 1)
            row[i] = sp.vstack([s.nlp.A_eq, L])
            row[self.N] = sp.vstack([sp.coo_matrix((s.nlp.n_eq, self.n_c)), -C])
            for j in range(self.N):
                if row[j] is None:
                    row[j] = sp.coo_matrix((s.nlp.n_eq + self.n_c, self.sc[j].nlp.n))
            rows.append(row)
        return sp.bmat(rows).tocsr()

    def _assemble_jacobian_ineq(self):
        blocks = [s.nlp.A_in for s in self.sc]
        J = sp.block_diag(blocks) if blocks else sp.coo_matrix((0, 0))
        return sp.hstack([J, sp.coo_matrix((J.shape[0], self.n_c))]).tocsr()

    def set_barrier_parameter(self, b):
        for s in self.sc:
            s.barrier = b

    # ---- KKT system in the reference's nested block format ----
    def evaluate_primal_dual_kkt_matrix(self, timer=None):
        """:1245-1285 (structure) + :1677-1681 (values)."""
        self.kkt_evals += 1
        N = self.N
        kkt = BlockMatrix(N + 1, N + 1)
        for i, (s, L, C) in enumerate(zip(self.sc, self.L, self.C)):
            n = s.nlp.n + s.nlp.n_eq + 2 * s.nlp.n_in
            sub = BlockMatrix(2, 2)
            sub.set_row_size(0, n); sub.set_col_size(0, n)
            sub.set_row_size(1, self.n_c); sub.set_col_size(1, self.n_c)
            row1 = BlockMatrix(1, 4)
            row1.set_row_size(0, self.n_c)
            for k, sz in enumerate((s.nlp.n, s.nlp.n_in, s.nlp.n_eq, s.nlp.n_in)):
                row1.set_col_size(k, sz)
            row1.set_block(0, 0, L)
            sub.set_block(0, 0, s.kkt())
            sub.set_block(1, 0, row1)
            sub.set_block(0, 1, row1.transpose())
            ptb = sp.identity(self.n_c, format="coo"); ptb.data.fill(0)
            sub.set_block(1, 1, ptb)
            kkt.set_block(i, i, sub)
            border = BlockMatrix(1, 2)
            border.set_col_size(0, n)
            border.set_block(0, 1, -C.transpose().tocoo())
            kkt.set_block(N, i, border)
            kkt.set_block(i, N, border.transpose())
        ptb = sp.identity(self.n_c, format="coo"); ptb.data.fill(0)
        kkt.set_block(N, N, ptb)
        return kkt

    def evaluate_primal_dual_kkt_rhs(self, timer=None):
        """:1683-1693."""
        N = self.N
        rhs = BlockVector(N + 1)
        last = np.zeros(self.n_c)
        for i, (s, L, C) in enumerate(zip(self.sc, self.L, self.C)):
            parts = s.rhs()
            parts[0] = parts[0] - L.T @ self.link_duals[i]
            inner = BlockVector(4)
            for k, p in enumerate(parts):
                inner.set_block(k, p)
            outer = BlockVector(2)
            outer.set_block(0, inner)
            outer.set_block(1, C @ self.z - L @ s.nlp.x)
            rhs.set_block(i, outer)
            last = last + C.T @ self.link_duals[i]
        rhs.set_block(N, last)
        return rhs

    def set_primal_dual_kkt_solution(self, sol):
        """:1695-1710 -- relies on the solver returning the nested structure of the right-hand side."""
        for i, s in enumerate(self.sc):
            inner = sol.get_block(i).get_block(0)
            s.set_solution([inner.get_block(k) for k in range(4)])
            self.d_link[i] = np.asarray(sol.get_block(i).get_block(1), dtype=float)
        self.d_z = np.asarray(sol.get_block(self.N), dtype=float)

    def get_delta_primals(self): return self._cat_x([s.d_x for s in self.sc], self.d_z)
    def get_delta_slacks(self): return np.concatenate([s.d_s for s in self.sc])
    def get_delta_duals_eq(self): return np.concatenate([np.concatenate([s.d_eq, d]) for s, d in zip(self.sc, self.d_link)])
    def get_delta_duals_ineq(self): return np.concatenate([s.d_in for s in self.sc])
    def get_delta_duals_primals_lb(self): return self._cat_x([s.d_zl() for s in self.sc], np.zeros(self.n_c))
    def get_delta_duals_primals_ub(self): return self._cat_x([s.d_zu() for s in self.sc], np.zeros(self.n_c))
    def get_delta_duals_slacks_lb(self): return np.concatenate([s.d_sl() for s in self.sc])
    def get_delta_duals_slacks_ub(self): return np.concatenate([s.d_su() for s in self.sc])

    def regularization_classes(self):
        """(per block, coupling): first-stage variables are the coupling primals (``:1755``)."""
        return {i: _block_classes(s, self.n_c) for i, s in enumerate(self.sc)}, np.full(self.n_c, 3, dtype=np.int8)

    def regularize_equality_gradient(self, kkt, coef, copy_kkt=True):
        """:1736-1745 with ``interface.py:590-608`` on every scenario."""
        if copy_kkt:
            kkt = kkt.copy()
        for i, s in enumerate(self.sc):
            inner = kkt.get_block(i, i).get_block(0, 0)
            inner.set_block(2, 2, coef * sp.identity(s.nlp.n_eq, format="coo"))
            inner.set_block(3, 3, coef * sp.identity(s.nlp.n_in, format="coo"))
            kkt.get_block(i, i).set_block(1, 1, coef * sp.identity(self.n_c, format="coo"))
        return kkt

    def regularize_hessian(self, kkt, coef, copy_kkt=True):
        """:1747-1757 with ``interface.py:610-619`` (``hess += coef*I`` changes the COO pattern)."""
        if copy_kkt:
            kkt = kkt.copy()
        for i, s in enumerate(self.sc):
            inner = kkt.get_block(i, i).get_block(0, 0)
            hess = inner.get_block(0, 0)
            hess = hess + coef * sp.identity(s.nlp.n, format="coo")
            inner.set_block(0, 0, hess)
        kkt.set_block(self.N, self.N, coef * sp.identity(self.n_c, format="coo"))
        return kkt


def _block_classes(s, n_link):
    """Diagonal classes of one diagonal block ``[x, s, lam_eq, lam_in, linking multipliers]`` for device-side
    regularisation (``parapint_b200/regularization.py``): what ``regularize_hessian`` / ``regularize_equality_gradient``
    touch (``interface.py:590-619``; linking rows ``sc_ip_interface.py:915-916,1744``)."""
    return np.concatenate([np.full(s.nlp.n, 1), np.zeros(s.nlp.n_in), np.full(s.nlp.n_eq + s.nlp.n_in + n_link, 2)]).astype(np.int8)


class DynamicInterface(StochasticInterface):
    """Restates ``DynamicSchurComplementInteriorPointInterface`` (``sc_ip_interface.py:22-1030``) on flat vectors:
    ``N`` time blocks, ``n_s`` states shared by neighbouring blocks through ``n_s (N-1)`` coupling variables ``z``.

    Per block ``i`` (``:143-177,359-475``): ``Lb_i`` selects the start states (``0 x n`` for block 0), ``Lf_i`` the end
    states (``0 x n`` for the last block); ``Cb_i`` / ``Cf_i`` select the coupling variables of the interface before /
    after the block.  Linking constraints ``Lb_i x_i - Cb_i z = 0`` and ``Lf_i x_i - Cf_i z = 0`` (``:716-739``).
    Equality duals per block are ordered ``[eq_i, backward link, forward link]`` (``:634-647``).

    KKT layout (``:274-357``): diagonal block ``[[KKT_i, Lb_i^T],[Lb_i, 0*I]]`` -- the *backward* multipliers live in
    the block --; coupling part ``[forward multipliers of every block ; z]`` with
    ``Q = [[0*I, -Cf],[-Cf^T, 0*I]]``; border of block ``i``: ``Lf_i`` into the rows of its forward multipliers
    and ``-Cb_i^T`` from its backward-multiplier columns into the rows of ``z``.  ``Q != 0`` and the border is a
    nested 2x2 block matrix: the structure BASELINE config 3 scales up."""

    def __init__(self, blocks, start_states, end_states):
        self.sc = [ScenarioInterface(b) for b in blocks]
        self.N = N = len(self.sc)
        self.n_s = n_s = len(start_states[0])
        self.n_c = n_s * (N - 1)                                   # :477-478
        self.nb = [0 if i == 0 else n_s for i in range(N)]        # backward-link rows per block
        self.nf = [0 if i == N - 1 else n_s for i in range(N)]    # forward-link rows per block

        def select(rows, idx, n):
            idx = np.asarray(idx[:rows], dtype=int)
            return sp.coo_matrix((np.ones(rows), (np.arange(rows), idx)), shape=(rows, n))

        self.Lb = [select(self.nb[i], start_states[i], b.n) for i, b in enumerate(blocks)]      # :418-446
        self.Lf = [select(self.nf[i], end_states[i], b.n) for i, b in enumerate(blocks)]        # :359-387
        self.Cb = [select(self.nb[i], n_s * (i - 1) + np.arange(n_s), self.n_c) for i in range(N)]   # :448-475
        self.Cf = [select(self.nf[i], n_s * i + np.arange(n_s), self.n_c) for i in range(N)]         # :389-416
        self.z = np.zeros(self.n_c)
        self.link_b = [np.zeros(k) for k in self.nb]
        self.link_f = [np.zeros(k) for k in self.nf]
        self.d_link_b = [np.zeros(k) for k in self.nb]
        self.d_link_f = [np.zeros(k) for k in self.nf]
        self.d_z = np.zeros(self.n_c)
        self.x_off = np.concatenate(([0], np.cumsum([s.nlp.n for s in self.sc])))
        self.s_off = np.concatenate(([0], np.cumsum([s.nlp.n_in for s in self.sc])))
        self.e_off = np.concatenate(([0], np.cumsum([s.nlp.n_eq + self.nb[i] + self.nf[i] for i, s in enumerate(self.sc)])))
        self.kkt_evals = 0

    def set_duals_eq(self, v):                                     # :659-677
        for i, s in enumerate(self.sc):
            seg = v[self.e_off[i]:self.e_off[i + 1]]
            ne = s.nlp.n_eq
            s.duals_eq = seg[:ne].copy()
            self.link_b[i] = seg[ne:ne + self.nb[i]].copy()
            self.link_f[i] = seg[ne + self.nb[i]:].copy()

    def get_duals_eq(self):
        return np.concatenate([np.concatenate([s.duals_eq, self.link_b[i], self.link_f[i]]) for i, s in enumerate(self.sc)])

    def evaluate_eq_constraints(self):                             # :716-739
        return np.concatenate([np.concatenate([s.nlp.eq(), self.Lb[i] @ s.nlp.x - self.Cb[i] @ self.z,
                                               self.Lf[i] @ s.nlp.x - self.Cf[i] @ self.z])
                               for i, s in enumerate(self.sc)])

    def _assemble_jacobian_eq(self):                               # :254-272,753-765
        rows = []
        for i, s in enumerate(self.sc):
            row = [sp.coo_matrix((s.nlp.n_eq + self.nb[i] + self.nf[i], t.nlp.n)) for t in self.sc] + [None]
            row[i] = sp.vstack([s.nlp.A_eq, self.Lb[i], self.Lf[i]])
            row[self.N] = sp.vstack([sp.coo_matrix((s.nlp.n_eq, self.n_c)), -self.Cb[i], -self.Cf[i]])
            rows.append(row)
        return sp.bmat(rows).tocsr()

    def evaluate_primal_dual_kkt_matrix(self, timer=None):
        """:274-357 (structure) + :839-843 (values)."""
        self.kkt_evals += 1
        N, n_s = self.N, self.n_s
        kkt = BlockMatrix(N + 1, N + 1)
        for i, s in enumerate(self.sc):
            n = s.nlp.n + s.nlp.n_eq + 2 * s.nlp.n_in
            sizes = (s.nlp.n, s.nlp.n_in, s.nlp.n_eq, s.nlp.n_in)
            sub = BlockMatrix(2, 2)
            sub.set_row_size(0, n); sub.set_col_size(0, n)
            sub.set_row_size(1, self.nb[i]); sub.set_col_size(1, self.nb[i])
            ptb = sp.identity(self.nb[i], format="coo"); ptb.data.fill(0)
            sub.set_block(1, 1, ptb)
            row1 = BlockMatrix(1, 4)
            row1.set_row_size(0, self.nb[i])
            for k, sz in enumerate(sizes):
                row1.set_col_size(k, sz)
            row1.set_block(0, 0, self.Lb[i])
            sub.set_block(1, 0, row1)
            sub.set_block(0, 1, row1.transpose())
            sub.set_block(0, 0, s.kkt())
            kkt.set_block(i, i, sub)
            # border (:318-333): [[forward-link rows of every block x KKT_i columns, .],[., -Cb_i^T]]
            border = BlockMatrix(2, 2)
            fwd = BlockMatrix(N, 4)
            for k, sz in enumerate(sizes):
                fwd.set_col_size(k, sz)
            for j in range(N):
                fwd.set_row_size(j, self.nf[j])
            fwd.set_block(i, 0, self.Lf[i])
            border.set_block(0, 0, fwd)
            border.set_block(1, 1, (-self.Cb[i].transpose()).tocoo())
            kkt.set_block(N, i, border)
            kkt.set_block(i, N, border.transpose())
        # bottom-right block (:335-357)
        Q = BlockMatrix(2, 2)
        sub = BlockMatrix(1, N)
        for j in range(N):
            sub.set_block(0, j, (-self.Cf[j].transpose()).tocoo())
        Q.set_block(1, 0, sub)
        Q.set_block(0, 1, sub.transpose())
        ptb = sp.identity(n_s * (N - 1), format="coo"); ptb.data.fill(0)
        Q.set_block(0, 0, ptb)
        ptb = sp.identity(self.n_c, format="coo"); ptb.data.fill(0)
        Q.set_block(1, 1, ptb)
        kkt.set_block(N, N, Q)
        return kkt

    def evaluate_primal_dual_kkt_rhs(self, timer=None):
        """:845-862."""
        N = self.N
        rhs = BlockVector(N + 1)
        fwd = BlockVector(N)
        last = np.zeros(self.n_c)
        for i, s in enumerate(self.sc):
            parts = s.rhs()
            parts[0] = parts[0] - (self.Lb[i].T @ self.link_b[i] + self.Lf[i].T @ self.link_f[i])
            inner = BlockVector(4)
            for k, p in enumerate(parts):
                inner.set_block(k, p)
            outer = BlockVector(2)
            outer.set_block(0, inner)
            outer.set_block(1, self.Cb[i] @ self.z - self.Lb[i] @ s.nlp.x)
            rhs.set_block(i, outer)
            fwd.set_block(i, self.Cf[i] @ self.z - self.Lf[i] @ s.nlp.x)
            last = last + self.Cb[i].T @ self.link_b[i] + self.Cf[i].T @ self.link_f[i]
        tail = BlockVector(2)
        tail.set_block(0, fwd)
        tail.set_block(1, last)
        rhs.set_block(N, tail)
        return rhs

    def set_primal_dual_kkt_solution(self, sol):
        """:864-877 -- relies on the solver returning the nested structure of the right-hand side."""
        for i, s in enumerate(self.sc):
            inner = sol.get_block(i).get_block(0)
            s.set_solution([inner.get_block(k) for k in range(4)])
            self.d_link_b[i] = np.asarray(sol.get_block(i).get_block(1), dtype=float)
            self.d_link_f[i] = np.asarray(sol.get_block(self.N).get_block(0).get_block(i), dtype=float)
        self.d_z = np.asarray(sol.get_block(self.N).get_block(1), dtype=float)

    def get_delta_duals_eq(self):
        return np.concatenate([np.concatenate([s.d_eq, self.d_link_b[i], self.d_link_f[i]]) for i, s in enumerate(self.sc)])

    def regularization_classes(self):
        """(per block, coupling): the coupling part is [forward multipliers (``:917-919``) ; z (``:929-931``)]."""
        return ({i: _block_classes(s, self.nb[i]) for i, s in enumerate(self.sc)},
                np.concatenate([np.full(self.n_s * (self.N - 1), 2), np.full(self.n_c, 3)]).astype(np.int8))

    def regularize_equality_gradient(self, kkt, coef, copy_kkt=True):
        """:903-920."""
        if copy_kkt:
            kkt = kkt.copy()
        for i, s in enumerate(self.sc):
            inner = kkt.get_block(i, i).get_block(0, 0)
            inner.set_block(2, 2, coef * sp.identity(s.nlp.n_eq, format="coo"))
            inner.set_block(3, 3, coef * sp.identity(s.nlp.n_in, format="coo"))
            kkt.get_block(i, i).set_block(1, 1, coef * sp.identity(self.nb[i], format="coo"))
        block = kkt.get_block(self.N, self.N)
        block.set_block(0, 0, coef * sp.identity(block.get_row_size(0), format="coo"))
        kkt.set_block(self.N, self.N, block)
        return kkt

    def regularize_hessian(self, kkt, coef, copy_kkt=True):
        """:922-933."""
        if copy_kkt:
            kkt = kkt.copy()
        for i, s in enumerate(self.sc):
            inner = kkt.get_block(i, i).get_block(0, 0)
            inner.set_block(0, 0, inner.get_block(0, 0) + coef * sp.identity(s.nlp.n, format="coo"))
        block = kkt.get_block(self.N, self.N)
        block.set_block(1, 1, coef * sp.identity(block.get_row_size(1), format="coo"))
        kkt.set_block(self.N, self.N, block)
        return kkt


def device_regularized(cls):
    """The same interface with the ``regularize_*`` methods replaced by the device-side variant
    (``parapint_b200.regularization.DeviceRegularizationMixin``): what a parapint user adds to switch it on."""
    from parapint_b200.regularization import DeviceRegularizationMixin
    return type("DeviceRegularized" + cls.__name__, (DeviceRegularizationMixin, cls), {})


# --------------------------------------------------------------------------------------------------
class IPOptions:
    """Defaults of ``interior_point.py:57-60,86-88,159-171``."""
    max_iter = 1000
    tol = 1e-8
    init_barrier_parameter = 0.1
    minimum_barrier_parameter = 1e-9
    barrier_decrease = 10
    use_inertia_correction = True
    error_scaling = 100
    bounds_relaxation_factor = 1e-8
    init_coef = 1e-8
    factor_increase = 10
    factor_decrease = 1.0 / 3.0
    max_coef = 1e9
    reallocation_factor = 2
    max_num_reallocations = 5


def process_init(x, lb, ub):
    """:761-788 (compression-matrix arithmetic written as masks)."""
    if np.any(ub - lb < 0):
        raise ValueError("Lower bounds for variables/inequalities should not be larger than upper bounds.")
    if np.any(ub - lb == 0):
        raise ValueError("Variables and inequalities should not have equal lower and upper bounds.")
    has_lb, has_ub = ~np.isneginf(lb), ~np.isinf(ub)
    out = (x >= ub) | (x <= lb)
    m = out & has_lb & ~has_ub
    x[m] = lb[m] + 1
    m = out & has_ub & ~has_lb
    x[m] = ub[m] - 1
    m = out & has_lb & has_ub
    x[m] = 0.5 * (lb[m] + ub[m])


def process_init_duals_lb(x, lb):
    x[x <= 0] = 1
    x[np.isneginf(lb)] = 0


def process_init_duals_ub(x, ub):
    x[x <= 0] = 1
    x[np.isinf(ub)] = 0


def _max_abs(v):
    return 0 if v.size == 0 else np.max(np.abs(v))


def check_convergence(interface, barrier, error_scaling):
    """:174-317."""
    slacks = interface.get_slacks()
    grad_obj = interface.get_obj_factor() * interface.evaluate_grad_objective()
    jac_eq, jac_in = interface.evaluate_jacobian_eq(), interface.evaluate_jacobian_ineq()
    eq_resid = interface.evaluate_eq_constraints()
    in_resid = interface.evaluate_ineq_constraints() - slacks
    x = interface.get_primals()
    lam_eq, lam_in = interface.get_duals_eq(), interface.get_duals_ineq()
    zl, zu = interface.get_duals_primals_lb(), interface.get_duals_primals_ub()
    sl, su = interface.get_duals_slacks_lb(), interface.get_duals_slacks_ub()
    xl, xu = interface.primals_lb(), interface.primals_ub()
    gl, gu = interface.ineq_lb(), interface.ineq_ub()
    xl_m, xu_m, gl_m, gu_m = xl.copy(), xu.copy(), gl.copy(), gu.copy()
    xl_m[np.isneginf(xl)] = 0; xu_m[np.isinf(xu)] = 0
    gl_m[np.isneginf(gl)] = 0; gu_m[np.isinf(gu)] = 0
    grad_lag_x = grad_obj + jac_eq.T @ lam_eq + jac_in.T @ lam_in - zl + zu
    grad_lag_s = -lam_in - sl + su
    r_xl = (x - xl_m) * zl - barrier; r_xu = (xu_m - x) * zu - barrier
    r_xl[np.isneginf(xl)] = 0; r_xu[np.isinf(xu)] = 0
    r_sl = (slacks - gl_m) * sl - barrier; r_su = (gu_m - slacks) * su - barrier
    r_sl[np.isneginf(gl)] = 0; r_su[np.isinf(gu)] = 0
    primal_inf = max(_max_abs(eq_resid), _max_abs(in_resid))
    dual_inf = max(np.max(np.abs(grad_lag_x)), _max_abs(grad_lag_s))
    compl_inf = max(_max_abs(r_xl), _max_abs(r_xu), _max_abs(r_sl), _max_abs(r_su))
    nbounds = np.isfinite(xl).sum() + np.isfinite(xu).sum() + np.isfinite(gl).sum() + np.isfinite(gu).sum()
    bsum = np.abs(zl).sum() + np.abs(zu).sum() + np.abs(sl).sum() + np.abs(su).sum()
    dual_scaling = (np.abs(lam_eq).sum() + np.abs(lam_in).sum() + bsum) / (lam_eq.size + lam_in.size + nbounds)
    dual_scaling = max(error_scaling, dual_scaling) / error_scaling
    compl_scaling = max(error_scaling, bsum / nbounds) / error_scaling
    return primal_inf, dual_inf / dual_scaling, compl_inf / compl_scaling


def _ftb_lb(tau, x, dx, xl):
    """:655-663."""
    d = dx.copy(); d[d == 0] = 1
    alpha = -tau * (x - xl) / d
    alpha[dx >= 0] = np.inf
    return 1 if alpha.size == 0 else min(alpha.min(), 1)


def _ftb_ub(tau, x, dx, xu):
    d = dx.copy(); d[d == 0] = 1
    alpha = tau * (xu - x) / d
    alpha[dx <= 0] = np.inf
    return 1 if alpha.size == 0 else min(alpha.min(), 1)


def fraction_to_the_boundary(interface, tau):
    """:677-758."""
    x, s = interface.get_primals(), interface.get_slacks()
    dx, ds = interface.get_delta_primals(), interface.get_delta_slacks()
    a_p = min(_ftb_lb(tau, x, dx, interface.primals_lb()), _ftb_ub(tau, x, dx, interface.primals_ub()),
              _ftb_lb(tau, s, ds, interface.ineq_lb()), _ftb_ub(tau, s, ds, interface.ineq_ub()))
    zl, zu = interface.get_duals_primals_lb(), interface.get_duals_primals_ub()
    sl, su = interface.get_duals_slacks_lb(), interface.get_duals_slacks_ub()
    a_d = min(_ftb_lb(tau, zl, interface.get_delta_duals_primals_lb(), np.zeros_like(zl)),
              _ftb_lb(tau, zu, interface.get_delta_duals_primals_ub(), np.zeros_like(zu)),
              _ftb_lb(tau, sl, interface.get_delta_duals_slacks_lb(), np.zeros_like(sl)),
              _ftb_lb(tau, su, interface.get_delta_duals_slacks_ub(), np.zeros_like(su)))
    return a_p, a_d


def try_factorization_and_reallocation(kkt, solver, opts, which):
    """:634-652 -- calls the plugin BY KEYWORD, as the reference does."""
    method = solver.do_numeric_factorization if which == "numeric" else solver.do_symbolic_factorization
    for count in range(opts.max_num_reallocations):
        res = method(matrix=kkt, raise_on_error=False, timer=None)
        status = res.status
        if status == LinearSolverStatus.not_enough_memory:
            solver.increase_memory_allocation(opts.reallocation_factor)
        else:
            break
    return status, count


def numeric_factorization(interface, kkt, solver, opts, inertia_coef, log):
    """Inertia-correction loop, :337-402."""
    status, _ = try_factorization_and_reallocation(kkt, solver, opts, "numeric")
    final = 0
    if status not in (LinearSolverStatus.successful, LinearSolverStatus.singular):
        raise RuntimeError("Could not factorize KKT system; linear solver status: " + str(status))
    neg = zero = None
    it = 0
    target = interface.n_eq_constraints() + interface.n_ineq_constraints()
    while final <= opts.max_coef:
        if status == LinearSolverStatus.successful:
            _, neg, zero = solver.get_inertia()
        else:
            neg = zero = None
        log.append(("reg", it, final, neg, zero, status.name))
        if neg == target and zero == 0 and status == LinearSolverStatus.successful:
            break
        if it == 0:
            kkt = kkt.copy()
        kkt = interface.regularize_equality_gradient(kkt=kkt, coef=-inertia_coef, copy_kkt=False)
        kkt = interface.regularize_hessian(kkt=kkt, coef=inertia_coef, copy_kkt=False)
        status, _ = try_factorization_and_reallocation(kkt, solver, opts, "numeric")
        final = inertia_coef
        inertia_coef *= opts.factor_increase
        it += 1
    if neg != target or zero != 0 or status != LinearSolverStatus.successful:
        raise RuntimeError("Exceeded maximum inertia correciton")
    return final


def ip_solve(interface, solver, opts=None, check_convergence=None, fraction_to_the_boundary=None):
    """:405-631.  Returns a dict with status, iteration count, objective history and the regularisation log.
    ``check_convergence`` / ``fraction_to_the_boundary``: replacements for the two module functions of the same names
    (the tests pass the device versions of ``parapint_b200.ipm_vectors`` and demand the identical trajectory)."""
    check_convergence = check_convergence or globals()["check_convergence"]
    fraction_to_the_boundary = fraction_to_the_boundary or globals()["fraction_to_the_boundary"]
    opts = opts or IPOptions()
    interface.set_bounds_relaxation_factor(opts.bounds_relaxation_factor)
    barrier = opts.init_barrier_parameter
    inertia_coef = opts.init_coef
    x = interface.init_primals().copy(); s = interface.init_slacks().copy()
    lam_eq = interface.init_duals_eq().copy(); lam_in = interface.init_duals_ineq().copy()
    zl = interface.init_duals_primals_lb().copy(); zu = interface.init_duals_primals_ub().copy()
    sl = interface.init_duals_slacks_lb().copy(); su = interface.init_duals_slacks_ub().copy()
    process_init(x, interface.primals_lb(), interface.primals_ub())
    process_init(s, interface.ineq_lb(), interface.ineq_ub())
    process_init_duals_lb(zl, interface.primals_lb()); process_init_duals_ub(zu, interface.primals_ub())
    process_init_duals_lb(sl, interface.ineq_lb()); process_init_duals_ub(su, interface.ineq_ub())
    interface.set_barrier_parameter(barrier)
    history, reglog = [], []
    status = "error"
    for it in range(opts.max_iter):
        interface.set_primals(x); interface.set_slacks(s)
        interface.set_duals_eq(lam_eq); interface.set_duals_ineq(lam_in)
        interface.set_duals_primals_lb(zl); interface.set_duals_primals_ub(zu)
        interface.set_duals_slacks_lb(sl); interface.set_duals_slacks_ub(su)
        p_inf, d_inf, c_inf = check_convergence(interface, 0, opts.error_scaling)
        history.append((it, interface.evaluate_objective(), p_inf, d_inf, c_inf, barrier))
        if max(p_inf, d_inf, c_inf) <= opts.tol:
            status = "optimal"
            break
        p_inf, d_inf, c_inf = check_convergence(interface, barrier, opts.error_scaling)
        if max(p_inf, d_inf, c_inf) <= opts.barrier_decrease * barrier:
            barrier = max(opts.minimum_barrier_parameter, min(0.5 * barrier, barrier ** 1.5))
        interface.set_barrier_parameter(barrier)
        kkt = interface.evaluate_primal_dual_kkt_matrix()
        rhs = interface.evaluate_primal_dual_kkt_rhs()
        if it == 0:
            st, _ = try_factorization_and_reallocation(kkt, solver, opts, "symbolic")
            if st != LinearSolverStatus.successful:
                raise RuntimeError("Could not factorize KKT system; linear solver status: " + str(st))
        used = numeric_factorization(interface, kkt, solver, opts, inertia_coef, reglog)
        inertia_coef = max(used * opts.factor_decrease, opts.init_coef)
        delta = solver.do_back_solve(rhs)
        interface.set_primal_dual_kkt_solution(delta)
        a_p, a_d = fraction_to_the_boundary(interface, 1 - barrier)
        x = x + a_p * interface.get_delta_primals()
        s = s + a_p * interface.get_delta_slacks()
        lam_eq = lam_eq + a_d * interface.get_delta_duals_eq()
        lam_in = lam_in + a_d * interface.get_delta_duals_ineq()
        zl = zl + a_d * interface.get_delta_duals_primals_lb()
        zu = zu + a_d * interface.get_delta_duals_primals_ub()
        sl = sl + a_d * interface.get_delta_duals_slacks_lb()
        su = su + a_d * interface.get_delta_duals_slacks_ub()
    return {"status": status, "iterations": len(history) - 1, "history": history, "reg": reglog,
            "objective": history[-1][1], "primals": interface.get_primals()}


# --------------------------------------------------------------------------------------------------
def farmer_scenarios():
    """The farmer two-stage LP of ``parapint/examples/stochastic.py:20-74`` with closed-form data.
    Variable order per scenario: acreage[3], sub-quota sold[3], super-quota sold[3], purchased[3]."""
    crops = ["WHEAT", "CORN", "SUGAR_BEETS"]
    total = 500.0
    price_quota = [100000.0, 100000.0, 6000.0]
    sub_price = [170.0, 150.0, 36.0]
    super_price = [0.0, 0.0, 10.0]
    feed = [200.0, 240.0, 0.0]
    purchase = [238.0, 210.0, 100000.0]
    plant = [150.0, 230.0, 260.0]
    yields = [[2.0, 2.4, 16.0], [2.5, 3.0, 20.0], [3.0, 3.6, 24.0]]
    probs = [0.3333, 0.3334, 0.3333]
    scen = []
    for y, p in zip(yields, probs):
        c = p * np.concatenate([plant, -np.asarray(sub_price), -np.asarray(super_price), purchase])
        lb = np.zeros(12)
        ub = np.concatenate([np.full(3, total), np.full(9, np.inf)])
        rows = []
        g_lb, g_ub = [], []
        r = np.zeros(12); r[0:3] = 1.0                       # total acreage <= 500
        rows.append(r); g_lb.append(-np.inf); g_ub.append(total)
        for i in range(3):                                   # feed requirement
            r = np.zeros(12); r[i] = y[i]; r[9 + i] = 1.0; r[3 + i] = -1.0; r[6 + i] = -1.0
            rows.append(r); g_lb.append(feed[i]); g_ub.append(np.inf)
        for i in range(3):                                   # limit amount sold
            r = np.zeros(12); r[3 + i] = 1.0; r[6 + i] = 1.0; r[i] = -y[i]
            rows.append(r); g_lb.append(-np.inf); g_ub.append(0.0)
        for i in range(3):                                   # quotas
            r = np.zeros(12); r[3 + i] = 1.0
            rows.append(r); g_lb.append(0.0); g_ub.append(price_quota[i])
        scen.append(QuadraticScenario(c, lb, ub, sp.coo_matrix(np.asarray(rows)), g_lb, g_ub))
    return scen, [[0, 1, 2]] * 3, crops


def random_stochastic_qp(seed, n_scen, n_x, n_eq, n_in, n_fs, negative_curvature=0.0):
    """A separable two-stage QP in the farmer mould, scaled up (SURVEY.md 8(d) "end-to-end inputs"): banded
    Hessians (optionally with some negative curvature so that the inertia-correction loop has work to do),
    banded equality rows, a few sparse inequality rows, box bounds on every variable; the first ``n_fs`` variables
    of every scenario are the first-stage variables."""
    scen = []
    for i in range(n_scen):
        rng = np.random.default_rng(7919 * seed + i)
        d = rng.uniform(0.5, 2.0, n_x)
        if negative_curvature > 0:
            pick = rng.random(n_x) < negative_curvature
            d[pick] = -rng.uniform(0.1, 0.5, pick.sum())
        off = rng.standard_normal(n_x - 1) * 0.2
        H = sp.diags([off, d, off], [-1, 0, 1]).tocoo()
        c = rng.standard_normal(n_x)
        A_eq = (sp.eye(n_eq, n_x, k=n_fs) + sp.diags([rng.standard_normal(n_eq) * 0.5], [n_fs + 1], shape=(n_eq, n_x))).tocoo()
        x_feas = rng.uniform(-1.0, 1.0, n_x)
        b_eq = A_eq @ x_feas
        A_in = (sp.random(n_in, n_x, density=3.0 / n_x, random_state=rng, data_rvs=rng.standard_normal)
                + sp.eye(n_in, n_x, k=n_x - n_in)).tocoo()
        g_mid = A_in @ x_feas
        scen.append(QuadraticScenario(c, np.full(n_x, -5.0), np.full(n_x, 5.0), A_in, g_mid - 2.0, g_mid + 3.0,
                                      A_eq=A_eq, b_eq=b_eq, H=H))
    return scen, [list(range(n_fs))] * n_scen


def dynamics_time_blocks(t0=0, delta_t=1, num_finite_elements=90, constant_control_duration=10, time_scale=0.1,
                         num_time_blocks=3):
    """The dynamics example of ``parapint/examples/dynamics.py:37-100,103-143`` with closed-form data:

        min  sum over finite elements of 0.5 dt [(x(t0) - sin(w t0) - 1)^2 + (x(t1) - sin(w t1) - 1)^2]   (trapezoid)
        s.t. x(t1) - (x(t0) + dt (p(t_p) - x(t1))) = 0        (implicit Euler of dx/dt = p - x),   p <= 2

    per time block; ``p`` is piecewise constant over ``constant_control_duration``.  Variable order per block:
    ``x`` at the block's ``nfe + 1`` time points, then its ``p`` values.  Returns the blocks, the start / end state
    indices and, per block, the times of its ``p`` variables (for the goldens of
    ``examples/tests/test_examples.py:47-57``)."""
    assert num_finite_elements % num_time_blocks == 0
    nfe = num_finite_elements // num_time_blocks
    assert constant_control_duration >= delta_t and constant_control_duration % delta_t == 0
    assert (nfe * delta_t) % constant_control_duration == 0
    blocks, starts, ends, p_times = [], [], [], []
    for b in range(num_time_blocks):
        bt0 = t0 + b * nfe * delta_t
        tx = [bt0 + k * delta_t for k in range(nfe + 1)]
        n_p = (nfe * delta_t) // constant_control_duration
        tp = [bt0 + k * constant_control_duration for k in range(n_p)]
        n = len(tx) + n_p
        h = np.zeros(n)
        c = np.zeros(n)
        const = 0.0
        for fe in range(nfe):                                  # dynamics.py:84-89
            for k in (fe, fe + 1):
                target = np.sin(time_scale * tx[k]) + 1.0
                h[k] += delta_t                                # 0.5 dt (x - target)^2 -> Hessian dt, gradient -dt*target
                c[k] += -delta_t * target
                const += 0.5 * delta_t * target * target
        rows, cols, vals = [], [], []
        for fe in range(nfe):                                  # dynamics.py:94-98
            ip = len(tx) + int(np.floor(fe / (constant_control_duration / delta_t)))
            rows += [fe, fe, fe]
            cols += [fe + 1, fe, ip]
            vals += [1.0 + delta_t, -1.0, -float(delta_t)]
        A_eq = sp.coo_matrix((vals, (rows, cols)), shape=(nfe, n))
        ub = np.full(n, np.inf)
        ub[len(tx):] = 2.0                                     # dynamics.py:80-81
        blk = QuadraticScenario(c, np.full(n, -np.inf), ub, None, np.zeros(0), np.zeros(0), A_eq=A_eq,
                                b_eq=np.zeros(nfe), H=sp.diags(h).tocoo())
        blk.const = const
        blocks.append(blk)
        starts.append([0])
        ends.append([len(tx) - 1])
        p_times.append(tp)
    return blocks, starts, ends, p_times


#: ``parapint/examples/tests/test_examples.py:47-57``: optimal controls p(t) of the dynamics example (7 places)
DYNAMICS_GOLDEN_P = {0: 1.6046242850486279, 10: 2.0, 20: 1.4792062911745605, 30: 0.5082444341496647,
                     40: -0.009859487375413882, 50: 0.40043954978583834, 60: 1.3619861771562247,
                     70: 1.99059057528143, 80: 1.7102013685364827}

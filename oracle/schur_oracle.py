"""numpy/scipy restatement of the reference's Schur-complement KKT solve.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows; the arithmetic order is kept (one leaf back-solve
per nonzero border row, ``S[:, r] -= A @ x``) so that results agree with the
reference to rounding.  The leaf is SciPy's SuperLU exactly as in reference
``parapint/linalg/scipy_interface.py:25-62``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
from scipy.sparse.linalg import splu

from parapint_b200.carriers import BlockVector, is_block_vector

SUCCESSFUL, NOT_ENOUGH_MEMORY, SINGULAR, ERROR, WARNING = 0, 1, 2, 3, 4


class LeafLU:
    """One diagonal block (or the coupling matrix) factorised by SuperLU.

    Follows ``scipy_interface.py:25-47`` (``splu(matrix.tocsc())``; status
    ``singular`` on "Factor is exactly singular"; optional inertia by dense
    ``eigvals`` thresholded at +-1e-8, ``:40-45``) and ``:49-62`` (solve keeps a
    ``BlockVector`` right-hand side's structure).
    """

    def __init__(self, compute_inertia=False, inertia_method="eigvals"):
        self.compute_inertia = compute_inertia
        self.inertia_method = inertia_method
        self.lu = None
        self.inertia = None
        self.status = None

    def factor(self, matrix):
        mat = matrix.tocsc() if not sp.isspmatrix_csc(matrix) else matrix
        try:
            self.lu = splu(mat)
            self.status = SUCCESSFUL
        except RuntimeError as err:
            self.lu = None
            self.status = SINGULAR if "Factor is exactly singular" in str(err) else ERROR
        if self.compute_inertia:
            self.inertia = dense_inertia(mat.toarray(), method=self.inertia_method)
        return self.status

    def solve(self, rhs):
        flat = rhs.flatten() if is_block_vector(rhs) else np.asarray(rhs, dtype=np.float64)
        sol = self.lu.solve(flat)
        if is_block_vector(rhs):
            out = rhs.copy_structure()
            out.copyfrom(sol)
            return out
        return sol


def dense_inertia(dense, method="eigvals", tol=1e-8):
    """(n_pos, n_neg, n_zero) of a dense symmetric matrix.

    ``eigvals``: the reference's rule, ``scipy_interface.py:40-45`` (general
    eigenvalues, ``> 1e-8`` positive, ``< -1e-8`` negative, rest zero).
    ``eigvalsh``: same thresholds on the symmetric eigen-solver (cheaper, real).
    ``ldl``: pivot signs of LAPACK ``dsytrf`` (Bunch-Kaufman), the convention of
    the MA27 / MUMPS leaves (``ma27_interface.py:201-203``,
    ``mumps_interface.py:122-126``): 1x1 pivots by sign, 2x2 pivots by the signs
    of their two eigenvalues; an exactly zero pivot counts as zero.
    """
    n = dense.shape[0]
    if n == 0:
        return 0, 0, 0
    if method == "eigvals":
        eig = sla.eigvals(dense)
    elif method == "eigvalsh":
        eig = sla.eigvalsh((dense + dense.T) * 0.5)
    elif method == "ldl":
        # LAPACK dsytrf called directly (scipy.linalg.ldl runs the same routine and then spends most of its time
        # rebuilding L and D in Python); lower storage: ipiv[k] > 0 is a 1x1 pivot, ipiv[k] = ipiv[k+1] < 0 a 2x2
        # pivot whose off-diagonal entry sits at (k+1, k)
        lwork, _ = sla.lapack.dsytrf_lwork(n, lower=1)      # blocked variant, as scipy.linalg.ldl asks for
        ldu, ipiv, info = sla.lapack.dsytrf(np.asfortranarray(dense), lower=1, lwork=max(int(lwork), 1))
        if info < 0:
            raise ValueError(f"dsytrf: illegal argument {-info}")
        diag = np.diagonal(ldu)
        first = np.zeros(n, dtype=bool)         # first column of every 2x2 pivot
        k = 0
        neg_piv = ipiv < 0
        while k < n:
            if neg_piv[k]:
                first[k] = True
                k += 2
            else:
                k += 1
        k2 = np.flatnonzero(first)
        single = np.ones(n, dtype=bool)
        single[k2] = False
        single[k2 + 1] = False
        d1 = diag[single]
        pos, neg, zero = int(np.count_nonzero(d1 > 0)), int(np.count_nonzero(d1 < 0)), int(np.count_nonzero(d1 == 0))
        if k2.size:
            off = ldu[k2 + 1, k2]
            blocks = np.empty((k2.size, 2, 2))
            blocks[:, 0, 0], blocks[:, 1, 1] = diag[k2], diag[k2 + 1]
            blocks[:, 0, 1] = blocks[:, 1, 0] = off
            ev = np.linalg.eigvalsh(blocks)     # one LAPACK call per block, batched by NumPy
            pos += int(np.count_nonzero(ev > 0))
            neg += int(np.count_nonzero(ev < 0))
            zero += int(np.count_nonzero(ev == 0))
        return pos, neg, zero
    else:
        raise ValueError(method)
    pos = int(np.count_nonzero(eig > tol))
    neg = int(np.count_nonzero(eig < -tol))
    return pos, neg, n - pos - neg


def _worst(status, sub_status):
    """``explicit_schur_complement.py:9-13``: a non-successful sub-status overwrites."""
    return status if sub_status == SUCCESSFUL else sub_status


class SchurOracle:
    """Serial algorithm of ``explicit_schur_complement.py:44-172``.

    ``local_blocks`` restricts the work to a rank's share, which together with
    :func:`solve_partitioned` restates ``mpi_explicit_schur_complement.py``
    (ownership ``:198-203``, local formation ``:312-333``, SUM-reduction
    ``:335-349``, replicated coupling factor ``:352-360``, solve ``:363-402``).
    """

    def __init__(self, compute_inertia=False, inertia_method="eigvals"):
        self.compute_inertia = compute_inertia
        self.inertia_method = inertia_method
        self.leaves = {}
        self.coupling_leaf = None
        self.kkt = None
        self.nblocks = 0
        self.local_blocks = []
        self.status = None

    def _new_leaf(self):
        return LeafLU(self.compute_inertia, self.inertia_method)

    # -- symbolic: ``explicit...:44-78`` (shape checks; SciPy leaves do nothing) --
    def symbolic(self, kkt, local_blocks=None):
        nbr, nbc = kkt.bshape
        if nbr != nbc:
            raise ValueError("The block matrix provided is not square.")
        self.nblocks = nbr - 1
        self.local_blocks = list(range(self.nblocks)) if local_blocks is None else list(local_blocks)
        self.leaves = {i: self._new_leaf() for i in self.local_blocks}
        self.coupling_leaf = self._new_leaf()
        self.status = SUCCESSFUL
        return self.status

    # -- numeric, local part: ``explicit...:99-121`` / ``mpi...:292-333`` --
    def local_contribution(self, kkt):
        """Factor the local diagonal blocks and return ``-sum_i A_i K_i^-1 A_i^T`` (dense)."""
        self.kkt = kkt
        N = self.nblocks
        m_c = kkt.get_row_size(N) if hasattr(kkt, "get_row_size") else kkt.get_block(N, N).shape[0]
        status = SUCCESSFUL
        for i in self.local_blocks:
            status = _worst(status, self.leaves[i].factor(kkt.get_block(i, i)))
            if status not in (SUCCESSFUL, WARNING):
                break
        self.status = status
        contrib = np.zeros((m_c, m_c))
        if status not in (SUCCESSFUL, WARNING):
            return status, contrib
        for i in self.local_blocks:
            A = kkt.get_block(N, i).tocsr()
            for r in range(A.shape[0]):
                if A.indptr[r + 1] == A.indptr[r]:
                    continue
                rhs = A[r, :].toarray()[0]
                x = self.leaves[i].solve(rhs)
                contrib[:, r] -= A.dot(x)
        return status, contrib

    # -- numeric, coupling part: ``explicit...:108,122-128`` / ``mpi...:347-358`` --
    def factor_coupling(self, kkt, contrib_sum):
        N = self.nblocks
        S = kkt.get_block(N, N).toarray() + contrib_sum
        self.S = S
        self.status = _worst(self.status, self.coupling_leaf.factor(sp.coo_matrix(S)))
        return self.status

    def numeric(self, kkt):
        status, contrib = self.local_contribution(kkt)
        if status not in (SUCCESSFUL, WARNING):
            return status
        return self.factor_coupling(kkt, contrib)

    # -- solve: ``explicit...:131-155`` / ``mpi...:363-402`` (no mutation of rhs) --
    def local_forward(self, rhs):
        N = self.nblocks
        r_c = np.zeros(self.kkt.get_block(N, N).shape[0])
        for i in self.local_blocks:
            A = self.kkt.get_block(N, i).tocsr()
            y = self.leaves[i].solve(rhs.get_block(i))
            r_c -= A.dot(y.flatten() if is_block_vector(y) else y)
        return r_c

    def solve_coupling(self, rhs, r_c_sum):
        N = self.nblocks
        top = rhs.get_block(N)
        flat = top.flatten() if is_block_vector(top) else np.asarray(top, dtype=np.float64)
        sol = self.coupling_leaf.solve(flat + r_c_sum)
        if is_block_vector(top):   # the leaf keeps the nested structure of its right-hand side (scipy_interface.py:57-60)
            out = top.copy_structure()
            out.copyfrom(sol)
            return out
        return sol

    def local_backward(self, rhs, x_c, out):
        N = self.nblocks
        for i in self.local_blocks:
            At = self.kkt.get_block(N, i).tocsr().transpose()
            xc_flat = x_c.flatten() if is_block_vector(x_c) else x_c   # explicit...:153 `coupling.flatten()`
            out.set_block(i, self.leaves[i].solve(rhs.get_block(i) - At.dot(xc_flat)))
        return out

    def solve(self, rhs):
        out = BlockVector(self.nblocks + 1)
        x_c = self.solve_coupling(rhs, self.local_forward(rhs))
        out.set_block(self.nblocks, x_c)
        return self.local_backward(rhs, x_c, out)

    # -- inertia: ``explicit...:157-172`` / ``mpi...:404-436`` --
    def local_inertia(self):
        tot = np.zeros(3, dtype=np.int64)
        for i in self.local_blocks:
            if self.leaves[i].inertia is None:
                raise RuntimeError("The intertia was not computed during do_numeric_factorization.")
            tot += np.asarray(self.leaves[i].inertia, dtype=np.int64)
        return tot

    def inertia(self):
        tot = self.local_inertia() + np.asarray(self.coupling_leaf.inertia, dtype=np.int64)
        return int(tot[0]), int(tot[1]), int(tot[2])


def owner_of(block, size):
    """Round-robin ownership, ``mpi_sc_ip_interface.py:14-19`` / perf ``utils.py:6-21``."""
    return block % size


def solve_partitioned(kkt, rhs, size, compute_inertia=False, inertia_method="eigvals"):
    """Emulate ``MPISchurComplementLinearSolver`` on ``size`` ranks in one process.

    Every rank forms its local ``S`` share; the shares are summed in rank order
    (the Allreduce of ``mpi...:343``), every rank factors the full ``S``; the
    coupling right-hand side is summed likewise (``:387``).  Returns
    ``(status, x, inertia)`` with ``x`` assembled from the owners' blocks.
    """
    N = kkt.bshape[0] - 1
    ranks = []
    for r in range(size):
        o = SchurOracle(compute_inertia, inertia_method)
        o.symbolic(kkt, [i for i in range(N) if owner_of(i, size) == r])
        ranks.append(o)
    parts = [o.local_contribution(kkt) for o in ranks]
    status = SUCCESSFUL
    for st, _ in parts:
        status = _worst(status, st)
    if status not in (SUCCESSFUL, WARNING):
        return status, None, None
    total = np.zeros_like(parts[0][1])
    for _, c in parts:
        total += c
    for o in ranks:
        status = _worst(status, o.factor_coupling(kkt, total))
    if status not in (SUCCESSFUL, WARNING):
        return status, None, None
    r_c = np.zeros(total.shape[0])
    for o in ranks:
        r_c += o.local_forward(rhs)
    out = BlockVector(N + 1)
    x_c = ranks[0].solve_coupling(rhs, r_c)
    out.set_block(N, x_c)
    for o in ranks:
        o.local_backward(rhs, x_c, out)
    inertia = None
    if compute_inertia:
        tot = sum(o.local_inertia() for o in ranks) + np.asarray(ranks[0].coupling_leaf.inertia, dtype=np.int64)
        inertia = (int(tot[0]), int(tot[1]), int(tot[2]))
    return status, out, inertia


def full_space_solve(kkt, rhs):
    """Undecomposed solve (``performance/schur_complement/main.py:84-86`` 'fs')."""
    lu = splu(sym_full(kkt).tocsc())
    return lu.solve(rhs.flatten())


def sym_full(kkt):
    """Whole symmetric matrix from a block matrix that may carry only the lower border."""
    N = kkt.bshape[0] - 1
    sizes = [kkt.get_block(i, i).shape[0] for i in range(N + 1)]
    off = np.concatenate(([0], np.cumsum(sizes)))
    rows, cols, vals = [], [], []

    def put(blk, r0, c0):
        c = blk.tocoo()
        rows.append(c.row.astype(np.int64) + r0)
        cols.append(c.col.astype(np.int64) + c0)
        vals.append(c.data.astype(np.float64))

    for i in range(N + 1):
        put(kkt.get_block(i, i), off[i], off[i])
    for i in range(N):
        A = kkt.get_block(N, i)
        put(A, off[N], off[i])
        put(A.transpose(), off[i], off[N])
    n = off[-1]
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()


class OraclePlugin:
    """The reference's serial ``SchurComplementLinearSolver`` + ``ScipyInterface(compute_inertia=True)`` leaves
    behind the four-method plugin surface (``base_linear_solver_interface.py:26-51``), for driving the flat IPM
    restatement (``oracle/ipm.py``) with the reference algorithm."""

    def __init__(self, inertia_method="eigvals"):
        from parapint_b200.interface import LinearSolverResults, LinearSolverStatus
        self._Results, self._Status = LinearSolverResults, LinearSolverStatus
        self.oracle = SchurOracle(compute_inertia=True, inertia_method=inertia_method)
        self.n_numeric = 0

    def _res(self, code):
        r = self._Results()
        r.status = self._Status(code)
        return r

    def do_symbolic_factorization(self, matrix, raise_on_error=True, timer=None):
        return self._res(self.oracle.symbolic(matrix))

    def do_numeric_factorization(self, matrix, raise_on_error=True, timer=None):
        self.n_numeric += 1
        return self._res(self.oracle.numeric(matrix))

    def do_back_solve(self, rhs):
        return self.oracle.solve(rhs)

    def get_inertia(self):
        return self.oracle.inertia()

    def increase_memory_allocation(self, factor):
        pass

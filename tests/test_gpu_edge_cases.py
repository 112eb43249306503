"""GPU parity on the corner cases the reference's inputs allow: ragged blocks, a block without
border entries, a 1 x 1 block, a single block, duplicates and explicit zeros in the COO leaves, a missing (None)
coupling diagonal, values that change between factorisations, many right-hand sides per factorisation.
Checked against dense NumPy on the assembled matrix: solution to 1e-10 (the systems are tiny and well conditioned
by construction), inertia against ``eigvalsh``."""
import numpy as np
import pytest
import scipy.sparse as sp

from parapint_b200 import B200SchurComplementLinearSolver, BlockMatrix, BlockVector, LinearSolverStatus
from tests.helpers import block_vector

pytestmark = pytest.mark.gpu


def _sym(rng, n, shift=0.0):
    M = rng.standard_normal((n, n))
    return M + M.T + np.diag(rng.standard_normal(n) * 3.0 + shift)


def _dense(kkt, sizes):
    """The whole symmetric matrix as the solver reads it: lower triangles of the diagonal blocks mirrored, the lower
    border standing for both borders, blocks that are not set meaning zero."""
    off = np.concatenate(([0], np.cumsum(sizes)))
    N = len(sizes) - 1
    K = np.zeros((off[-1], off[-1]))
    for i in range(N + 1):
        blk = kkt.get_block(i, i)
        if blk is not None:
            d = np.tril(blk.toarray())
            K[off[i]:off[i + 1], off[i]:off[i + 1]] = d + np.tril(d, -1).T
        if i < N and kkt.get_block(N, i) is not None:
            A = kkt.get_block(N, i).toarray()
            K[off[N]:, off[i]:off[i + 1]] = A
            K[off[i]:off[i + 1], off[N]:] = A.T
    return K


def _check(kkt, sizes, rng, solver=None, n_rhs=1):
    s = solver or B200SchurComplementLinearSolver()
    if solver is None:
        assert s.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    K = _dense(kkt, sizes)
    ev = np.linalg.eigvalsh(K)
    assert np.min(np.abs(ev)) > 1e-6, "test matrix too close to singular"
    assert s.get_inertia() == (int((ev > 0).sum()), int((ev < 0).sum()), 0)
    for _ in range(n_rhs):
        b = rng.standard_normal(sum(sizes))
        x = s.do_back_solve(block_vector(b, sizes)).flatten()
        x_ref = np.linalg.solve(K, b)
        assert np.linalg.norm(x - x_ref) <= 1e-10 * np.linalg.norm(x_ref)
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b)
    return s


def test_ragged_blocks_empty_border_and_one_by_one_block():
    rng = np.random.default_rng(0)
    sizes = [7, 1, 23, 4, 3]            # four blocks of different orders (one of them 1 x 1), three coupling variables
    nb, m_c = len(sizes) - 1, sizes[-1]
    kkt = BlockMatrix(nb + 1, nb + 1)
    for i, n in enumerate(sizes[:-1]):
        kkt.set_block(i, i, sp.coo_matrix(_sym(rng, n)))
        A = rng.standard_normal((m_c, n))
        if i == 2:
            A[:] = 0.0                   # this block does not touch the coupling variables at all (stored, but empty)
        if i == 3:
            A[1] = 0.0                   # and this one only two of the three
        kkt.set_block(nb, i, sp.coo_matrix(A))
    kkt.set_block(nb, nb, sp.coo_matrix(_sym(rng, m_c)))
    _check(kkt, sizes, rng, n_rhs=3)


def test_border_block_absent():
    """A border block that is not set at all (``get_block(N, i) is None``) is an empty border."""
    rng = np.random.default_rng(1)
    sizes = [5, 6, 2]
    kkt = BlockMatrix(3, 3)
    kkt.set_block(0, 0, sp.coo_matrix(_sym(rng, 5)))
    kkt.set_block(1, 1, sp.coo_matrix(_sym(rng, 6)))
    kkt.set_block(2, 0, sp.coo_matrix(rng.standard_normal((2, 5))))
    kkt.set_block(2, 2, sp.coo_matrix(_sym(rng, 2)))
    for i in range(3):
        kkt.set_row_size(i, sizes[i])
        kkt.set_col_size(i, sizes[i])
    s = B200SchurComplementLinearSolver()
    res = s.do_symbolic_factorization(kkt, raise_on_error=False)
    if res.status != LinearSolverStatus.successful:
        pytest.skip("an unset border block is rejected at the symbolic phase (the reference requires every (N, i) block)")
    _check(kkt, sizes, rng, solver=s)


def test_single_block_single_coupling_variable():
    rng = np.random.default_rng(2)
    sizes = [9, 1]
    kkt = BlockMatrix(2, 2)
    kkt.set_block(0, 0, sp.coo_matrix(_sym(rng, 9)))
    kkt.set_block(1, 0, sp.coo_matrix(rng.standard_normal((1, 9))))
    kkt.set_block(1, 1, sp.coo_matrix(np.array([[-2.5]])))
    _check(kkt, sizes, rng)


def test_duplicates_and_explicit_zeros_in_coo_leaves():
    """COO leaves may repeat an entry (the values add up) and store explicit zeros; both triangles of
    K_i are stored, only the lower one is read."""
    rng = np.random.default_rng(3)
    n, m_c = 12, 3
    sizes = [n, n, m_c]
    kkt = BlockMatrix(3, 3)
    for i in range(2):
        K = _sym(rng, n)
        r, c = np.nonzero(np.ones((n, n)))
        v = K[r, c]
        # split every diagonal entry in two duplicates, append explicit zeros on a few positions
        dr = np.arange(n)
        rows = np.concatenate([r, dr, [3, 7, 7]])
        cols = np.concatenate([c, dr, [1, 2, 2]])
        vals = np.concatenate([v, np.zeros(n), [0.0, 0.0, 0.0]])
        vals[: n * n][r == c] *= 0.5
        vals[n * n: n * n + n] = 0.5 * np.diag(K)
        kkt.set_block(i, i, sp.coo_matrix((vals, (rows, cols)), shape=(n, n)))
        A = rng.standard_normal((m_c, n))
        ar, ac = np.nonzero(np.ones((m_c, n)))
        kkt.set_block(2, i, sp.coo_matrix((np.concatenate([A[ar, ac] * 0.25, A[ar, ac] * 0.75]),
                                           (np.concatenate([ar, ar]), np.concatenate([ac, ac]))), shape=(m_c, n)))
    kkt.set_block(2, 2, sp.coo_matrix(_sym(rng, m_c)))
    _check(kkt, sizes, rng)


def test_coupling_diagonal_absent_means_zero():
    """Q not set (``None``): the stochastic layout has no coupling Hessian (sc_ip_interface.py:1282-1284 stores a zero
    block; a missing one must mean the same)."""
    rng = np.random.default_rng(4)
    n, m_c = 10, 2
    sizes = [n, n, n, m_c]
    kkt = BlockMatrix(4, 4)
    for i in range(3):
        kkt.set_block(i, i, sp.coo_matrix(_sym(rng, n, shift=8.0)))
        kkt.set_block(3, i, sp.coo_matrix(rng.standard_normal((m_c, n))))
    kkt.set_row_size(3, m_c)
    kkt.set_col_size(3, m_c)
    s = B200SchurComplementLinearSolver()
    res = s.do_symbolic_factorization(kkt, raise_on_error=False)
    if res.status != LinearSolverStatus.successful:
        pytest.skip("a missing coupling block is rejected at the symbolic phase")
    _check(kkt, sizes, rng, solver=s)


def test_values_change_between_factorisations_same_pattern():
    """Symbolic once, numeric many times with new values in the same leaves (what ip_solve does), several solves each."""
    rng = np.random.default_rng(5)
    sizes = [30, 30, 30, 4]
    pattern = [(_sym(rng, 30) != 0) for _ in range(3)]
    kkt = BlockMatrix(4, 4)
    leaves = []
    for i in range(3):
        K = sp.coo_matrix(_sym(rng, 30))
        A = sp.coo_matrix(rng.standard_normal((4, 30)))
        kkt.set_block(i, i, K)
        kkt.set_block(3, i, A)
        leaves += [K, A]
    Q = sp.coo_matrix(_sym(rng, 4))
    kkt.set_block(3, 3, Q)
    s = _check(kkt, sizes, rng, n_rhs=2)
    for _ in range(3):
        for i in range(3):
            Kd = _sym(rng, 30)
            leaves[2 * i].data[:] = Kd[leaves[2 * i].row, leaves[2 * i].col]       # in place: same objects, new values
            leaves[2 * i + 1].data[:] = rng.standard_normal(leaves[2 * i + 1].data.size)
        Qd = _sym(rng, 4)
        Q.data[:] = Qd[Q.row, Q.col]
        _check(kkt, sizes, rng, solver=s, n_rhs=2)
    assert s.symbolic_calls == 1

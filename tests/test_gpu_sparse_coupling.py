"""GPU parity of the SPARSE coupling path (block-tridiagonal S of time-decomposed problems, BASELINE config 3 in
miniature): the Schur complement is kept as the values of its pattern (mpi_explicit_schur_complement.py:228-255,
312-345) and factorised level by level; results must equal the reference algorithm's, which factorises the same S
with a sparse leaf."""
import numpy as np
import pytest

from oracle.schur_oracle import SchurOracle, dense_inertia, sym_full
from parapint_b200 import B200SchurComplementLinearSolver, LinearSolverStatus
from tests.helpers import block_vector, dynamic_ipm_system

pytestmark = pytest.mark.gpu


def _solve(kkt, rhs, **kw):
    s = B200SchurComplementLinearSolver(**kw)
    assert s.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    return s, s.do_back_solve(rhs)


def _rel_residual(kkt, x, rhs):
    K = sym_full(kkt)
    b = rhs.flatten()
    return np.linalg.norm(K @ x.flatten() - b) / np.linalg.norm(b)


@pytest.mark.parametrize("seed,N,n_x,n_eq,n_in,n_s", [(0, 12, 60, 30, 5, 4), (1, 33, 90, 40, 8, 6), (2, 64, 48, 20, 4, 3)])
def test_dynamic_small_vs_dense_and_oracle(seed, N, n_x, n_eq, n_in, n_s):
    kkt, sizes = dynamic_ipm_system(seed, N, n_x, n_eq, n_in, n_s)
    rng = np.random.default_rng(seed)
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs, options={"coupling_min_sparse": 16})
    cs = s.backend.coupling_stats()
    assert cs["levels"] >= 1 and cs["schur_size"] < cs["m_c"] ** 2 // 2, cs
    dense = sym_full(kkt).toarray()
    x_ref = np.linalg.solve(dense, rhs.flatten())
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == dense_inertia(dense, "ldl") == dense_inertia(dense, "eigvalsh")
    # the dense-S path of the same library gives the same answer
    s2, x2 = _solve(kkt, rhs, options={"coupling_min_sparse": 10 ** 6})
    assert s2.backend.coupling_stats()["levels"] == 0
    assert s2.get_inertia() == s.get_inertia()
    assert np.linalg.norm(x2.flatten() - x.flatten()) / np.linalg.norm(x_ref) <= 1e-9
    # refactor + re-solve reuse, a different right-hand side
    rhs2 = block_vector(rng.standard_normal(sum(sizes)), sizes)
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    xb = s.do_back_solve(rhs2)
    assert _rel_residual(kkt, xb, rhs2) <= 1e-10


def test_dynamic_default_thresholds_vs_oracle():
    """m_c = 2 * 50 * 15 = 1500 with the default thresholds: sparse S, interfaces of 50 states as in config 3."""
    N, n_s = 16, 50
    kkt, sizes = dynamic_ipm_system(4, N, 500, 380, 20, n_s)
    rng = np.random.default_rng(4)
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs)
    cs = s.backend.coupling_stats()
    assert cs["levels"] >= 1 and cs["m_c"] == 2 * n_s * (N - 1)
    o = SchurOracle()
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x_ref = o.solve(rhs).flatten()
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    # inertia: Haynsworth additivity with LAPACK on every block and on the (dense) S the oracle formed
    tot = np.zeros(3, dtype=np.int64)
    for i in range(N):
        tot += np.asarray(dense_inertia(kkt.get_block(i, i).toarray(), "ldl"), dtype=np.int64)
    tot += np.asarray(dense_inertia(0.5 * (o.S + o.S.T), "ldl"), dtype=np.int64)   # S as the oracle formed it
    assert s.get_inertia() == tuple(int(v) for v in tot)


def test_singular_coupling_system_reports_singular():
    """A singular S (two identical interface rows) must come back as `singular` through the levels, not as garbage."""
    kkt, sizes = dynamic_ipm_system(5, 12, 60, 30, 5, 4)
    # make the coupling rows of the last interface duplicate: zero the -I of Q for one coupling variable
    Q = kkt.get_block(12, 12).tocoo()
    data = Q.data.copy()
    nf = 4 * 11
    kill = np.where(((Q.row == nf + 3) & (Q.col == 3)) | ((Q.row == 3) & (Q.col == nf + 3)))[0]
    data[kill] = 0.0
    import scipy.sparse as sp
    kkt.set_block(12, 12, sp.coo_matrix((data, (Q.row, Q.col)), shape=Q.shape))
    # ... and the border entry that couples the same variable into block 1
    A = kkt.get_block(12, 1).tocoo()
    d = A.data.copy()
    d[A.row == nf + 3] = 0.0
    kkt.set_block(12, 1, sp.coo_matrix((d, (A.row, A.col)), shape=A.shape))
    dense = sym_full(kkt).toarray()
    assert np.linalg.matrix_rank(dense) < dense.shape[0]
    s = B200SchurComplementLinearSolver(options={"coupling_min_sparse": 16})
    s.do_symbolic_factorization(kkt)
    res = s.do_numeric_factorization(kkt, raise_on_error=False)
    assert res.status == LinearSolverStatus.singular


def test_dynamics_example_through_sparse_coupling():
    """parapint's dynamics example with 30 time blocks, coupling system held sparse: same trajectory as the oracle."""
    from oracle.schur_oracle import OraclePlugin
    from tests.test_dynamics import _run
    kw = dict(num_finite_elements=300, num_time_blocks=30)
    _, ref, p_ref = _run(OraclePlugin(), **kw)
    solver = B200SchurComplementLinearSolver(options={"coupling_min_sparse": 8})
    _, out, p = _run(solver, **kw)
    assert solver.backend.coupling_stats()["levels"] >= 1
    assert ref["status"] == out["status"] == "optimal" and out["iterations"] == ref["iterations"]
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    assert max(abs(p[t] - p_ref[t]) for t in p_ref) < 1e-8

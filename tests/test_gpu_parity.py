"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden fixtures."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.kkt_generator import EstimationModel
from oracle.schur_oracle import SchurOracle, dense_inertia, sym_full
from parapint_b200 import B200SchurComplementLinearSolver, LinearSolverStatus
from tests.helpers import block_vector, bordered_from_dense, random_bordered

pytestmark = pytest.mark.gpu


def _solve(kkt, rhs, **kw):
    s = B200SchurComplementLinearSolver(**kw)
    assert s.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    res = s.do_numeric_factorization(kkt)
    assert res.status == LinearSolverStatus.successful
    return s, s.do_back_solve(rhs)


def _rel_residual(kkt, x, rhs):
    K = sym_full(kkt)
    b = rhs.flatten()
    return np.linalg.norm(K @ x.flatten() - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", ["sym", "sym_q"])
def test_known_answer_8x8(known_answers, name):
    """reference test_explicit_schur_complement.py:13-55 / test_mpi_...:19-115 (symmetrised blocks)."""
    dense = known_answers[f"kat_{name}_dense"]
    rhs = block_vector(known_answers[f"kat_{name}_rhs"], [2, 2, 2, 2])
    kkt = bordered_from_dense(dense, [2, 2, 2, 2])
    s, x = _solve(kkt, rhs)
    assert np.allclose(x.flatten(), known_answers[f"kat_{name}_x"], rtol=1e-12, atol=1e-13)
    assert s.get_inertia() == tuple(known_answers[f"kat_{name}_inertia"])
    # refactor + re-solve reuse (test_mpi_explicit_schur_complement.py:113-115)
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    assert np.allclose(s.do_back_solve(rhs).flatten(), known_answers[f"kat_{name}_x"], rtol=1e-12, atol=1e-13)


def test_leaf_3x3(known_answers):
    """reference test_linear_solvers.py:63-79 through a 1-block system with an empty coupling row."""
    dense = np.zeros((4, 4))
    dense[:3, :3] = known_answers["leaf_dense"]
    dense[3, 3] = 1.0
    kkt = bordered_from_dense(dense, [3, 1])
    s = B200SchurComplementLinearSolver()
    zero = bordered_from_dense(dense, [3, 1])
    assert s.do_symbolic_factorization(zero).status == LinearSolverStatus.successful
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    for r, x in zip(known_answers["leaf_rhs"], known_answers["leaf_x"]):
        sol = s.do_back_solve(block_vector(np.concatenate([r, [2.0]]), [3, 1]))
        assert np.allclose(sol.flatten(), np.concatenate([x, [2.0]]), rtol=1e-12)
    pos, neg, zero_ = known_answers["leaf_inertia"]
    assert s.get_inertia() == (pos + 1, neg, zero_)


@pytest.mark.parametrize("tag", ["g_3_20_2_5", "g_4_60_3_10"])
def test_generator_small_vs_reference(generator_golden, tag):
    g = generator_golden
    args = tuple(int(v) for v in g[f"{tag}_args"])
    m = EstimationModel(*args)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s, x = _solve(kkt, rhs)
    ref = g[f"{tag}_x"]
    assert np.linalg.norm(x.flatten() - ref) / np.linalg.norm(ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == tuple(g[f"{tag}_inertia"]) == m.expected_inertia()
    assert abs(m.check_result(x) - g[f"{tag}_max_err"]) <= 1e-8


def test_generator_reference_golden():
    """reference examples/tests/test_examples.py:76-99: max_err 0.3163456780448639 (7 places)."""
    m = EstimationModel(3, 500, 12, 10)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s, x = _solve(kkt, rhs)
    assert abs(m.check_result(x) - 0.3163456780448639) < 5e-8
    assert s.get_inertia() == m.expected_inertia()
    assert _rel_residual(kkt, x, rhs) <= 1e-10


def test_config2_vs_oracle(generator_golden):
    """BASELINE config 2: 64 blocks x 2000 rows x 50 coupling; full size, against the live oracle."""
    m = EstimationModel(64, 150, 6, 50)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s, x = _solve(kkt, rhs)
    o = SchurOracle()
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x_ref = o.solve(rhs).flatten()
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == m.expected_inertia()
    assert abs(m.check_result(x) - generator_golden["g_64_150_6_50_max_err"]) <= 1e-8
    assert np.allclose(np.asarray(x.get_block(64)), generator_golden["g_64_150_6_50_xc"], rtol=1e-8)


@pytest.mark.parametrize("seed,n_blocks,n,m_c,rows", [(0, 3, 37, 5, None), (1, 5, 130, 9, 4), (2, 2, 300, 70, 33),
                                                      (3, 4, [17, 64, 129, 200], 12, 7)])
def test_random_general_border(seed, n_blocks, n, m_c, rows):
    """General (non-selection) borders, ragged block sizes, partial border rows; vs oracle + dense."""
    rng = np.random.default_rng(seed)
    kkt = random_bordered(rng, n_blocks, n, m_c, density=0.08, border_nnz_rows=rows)
    sizes = [kkt.get_block(i, i).shape[0] for i in range(n_blocks + 1)]
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs)
    dense = sym_full(kkt).toarray()
    x_ref = np.linalg.solve(dense, rhs.flatten())
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == dense_inertia(dense, "eigvalsh")
    o = SchurOracle(compute_inertia=True, inertia_method="eigvalsh")
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    assert np.linalg.norm(x.flatten() - o.solve(rhs).flatten()) / np.linalg.norm(x_ref) <= 1e-8
    assert s.get_inertia() == o.inertia()


def test_singular_block_reports_singular():
    dense = np.zeros((5, 5))
    dense[:2, :2] = [[1.0, 2.0], [2.0, 4.0]]  # rank 1
    dense[2:4, 2:4] = np.eye(2)
    dense[4, 4] = 1.0
    dense[4, 0] = dense[0, 4] = 1.0
    kkt = bordered_from_dense(dense, [2, 2, 1])
    s = B200SchurComplementLinearSolver()
    s.do_symbolic_factorization(kkt)
    res = s.do_numeric_factorization(kkt, raise_on_error=False)
    assert res.status == LinearSolverStatus.singular
    with pytest.raises(RuntimeError):
        s.do_numeric_factorization(kkt, raise_on_error=True)
    with pytest.raises(RuntimeError):
        s.get_inertia()


def test_non_square_rejected():
    from parapint_b200 import BlockMatrix
    bad = BlockMatrix(2, 3)
    with pytest.raises(ValueError):
        B200SchurComplementLinearSolver().do_symbolic_factorization(bad)


@pytest.mark.parametrize("seed,n_blocks,n_x,n_eq,n_in,n_fs", [(0, 4, 300, 240, 30, 10), (1, 3, 600, 500, 20, 25),
                                                               (2, 6, 220, 100, 60, 8)])
def test_ipm_shaped_stochastic_kkt(seed, n_blocks, n_x, n_eq, n_in, n_fs):
    """Family P: primal-dual KKT blocks with zero (2,2) blocks and barrier diagonals spread over
    eight orders of magnitude -- exercises 2x2 pivots and delayed pivots of the multifrontal path."""
    from tests.helpers import stochastic_ipm_system
    kkt, sizes = stochastic_ipm_system(seed, n_blocks, n_x, n_eq, n_in, n_fs)
    rng = np.random.default_rng(seed)
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs)
    stats = s.backend.plan_stats(0)
    assert stats["supernodes"] > 0 and not stats["fell_back_dense"]
    dense = sym_full(kkt).toarray()
    x_ref = np.linalg.solve(dense, rhs.flatten())
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    expect = dense_inertia(dense, "ldl")
    assert s.get_inertia() == expect == dense_inertia(dense, "eigvalsh")  # what interior_point.py:379 tests


def test_delayed_pivot_overflow_falls_back_to_dense():
    """With no delayed-pivot capacity the sparse path must notice and redo the block densely, not return garbage."""
    from tests.helpers import stochastic_ipm_system
    kkt, sizes = stochastic_ipm_system(3, 2, 300, 240, 30, 10)
    rng = np.random.default_rng(3)
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs, options={"sparse_dmax": 0})
    dense = sym_full(kkt).toarray()
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == dense_inertia(dense, "ldl")
    st = s.backend.plan_stats(0)
    assert st["fell_back_dense"] == 1 or st["delayed_to_root"] == 0


@pytest.mark.parametrize("options", [{}, {"panel_onchip": 0}, {"cluster_panel": 0}])
def test_large_dense_fronts(options):
    """Dense blocks tall enough for the cluster panel kernel (and, with cluster_panel=0, the single-CTA one):
    both must make the same pivot choices, so inertia and solution agree with LAPACK."""
    rng = np.random.default_rng(11)
    n, m_c, nb = 1300, 40, 2
    kkt = random_bordered(rng, nb, n, m_c, density=0.6, border_nnz_rows=25)
    sizes = [n] * nb + [m_c]
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s, x = _solve(kkt, rhs, options=options)
    assert s.backend.plan_stats(0)["supernodes"] == 0  # dense plan
    dense = sym_full(kkt).toarray()
    x_ref = np.linalg.solve(dense, rhs.flatten())
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    assert s.get_inertia() == dense_inertia(dense, "ldl")


def test_iterative_refinement():
    """The device residual b - K x (computed from the values that were factorised) agrees with numpy's, and
    refinement steps with a deliberately sloppy factorization (tiny pivot threshold => large element growth)
    bring the residual back under the bar."""
    from tests.helpers import stochastic_ipm_system
    kkt, sizes = stochastic_ipm_system(5, 3, 400, 300, 40, 12)
    rng = np.random.default_rng(5)
    rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    s = B200SchurComplementLinearSolver(options={"pivot_threshold": 1e-9}, max_refine=0)
    s.do_symbolic_factorization(kkt)
    s.do_numeric_factorization(kkt)
    raw = _rel_residual(kkt, s.do_back_solve(rhs), rhs)
    assert s.refine_steps == 0 and s.last_residual is None
    s.max_refine = 4
    x = s.do_back_solve(rhs)
    fin = _rel_residual(kkt, x, rhs)
    assert fin <= 1e-10 and fin <= 2 * raw
    assert s.last_residual is not None and 0.3 * fin <= s.last_residual <= 3 * fin + 1e-16
    if raw > s.refine_tol:
        assert s.refine_steps >= 1
    # accurate factorization: the check costs no correction solve
    s2, x2 = _solve(kkt, rhs)
    assert s2.refine_steps == 0 and s2.last_residual <= s2.refine_tol
    assert np.linalg.norm(x2.flatten() - x.flatten()) / np.linalg.norm(x2.flatten()) <= 1e-8


@pytest.mark.parametrize("shape,options", [((2, 300, 4, 200), {}), ((2, 800, 3, 800), {}),
                                           ((2, 800, 3, 800), {"panel_onchip": 0}),
                                           ((2, 800, 3, 800), {"cluster_panel": 0}),
                                           ((2, 800, 3, 800), {"update_strip": 2}),
                                           ((2, 800, 3, 800), {"update_strip": 5, "overlap_groups": 1}),
                                           ((2, 800, 3, 800), {"update_strip": 8}),
                                           ((2, 800, 3, 800), {"update_strip": 0})])
def test_wide_border_sparse_blocks(shape, options):
    """Config-5-shaped blocks (family G with a wide border): sparse subtree + a dense root front of
    `root columns + border rows` factorised by the panel kernel (single-CTA and, for the taller one, thread-block
    clusters with the panel's rows of L on chip or re-read from L2) and the DMMA update (one tile per CTA, or the
    software-pipelined kernel with strips of 2 / 5 / 8 tiles per CTA), with the root's pivot count set on the device.  Checked against the closed-form
    inertia, the residual bar, and the reference algorithm (oracle) on the same system."""
    m = EstimationModel(*shape)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s, x = _solve(kkt, rhs, options=options)
    st = s.backend.plan_stats(0)
    assert st["supernodes"] > 0 and not st["fell_back_dense"] and st["root_cols"] + shape[3] > 164
    assert s.get_inertia() == m.expected_inertia()
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    o = SchurOracle()
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x_ref = o.solve(rhs).flatten()
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8


@pytest.mark.parametrize("options", [{"defer_status": 0, "auto_residual": 0}, {"small_front": 0}, {"sparse": 0}])
def test_alternative_control_paths(options):
    """The synchronous status path several ranks fall back to (every phase reports its own status), the dense
    panel / interchange / update launch sequence for small fronts, and whole blocks as dense fronts must give the
    same answers as the default single-rank fast paths."""
    m = EstimationModel(6, 40, 3, 8)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s0, x0 = _solve(kkt, rhs)
    s1, x1 = _solve(kkt, rhs, options=options)
    assert s1.get_inertia() == s0.get_inertia() == m.expected_inertia()
    assert _rel_residual(kkt, x1, rhs) <= 1e-10
    assert np.linalg.norm(x1.flatten() - x0.flatten()) / np.linalg.norm(x0.flatten()) <= 1e-10
    # a singular block is reported by both status paths
    dense = np.zeros((5, 5))
    dense[:2, :2] = [[1.0, 2.0], [2.0, 4.0]]  # rank 1
    dense[2:4, 2:4] = np.eye(2)
    dense[4, 4] = 1.0
    dense[4, 0] = dense[0, 4] = 1.0
    bad = bordered_from_dense(dense, [2, 2, 1])
    s = B200SchurComplementLinearSolver(options=options)
    s.do_symbolic_factorization(bad)
    assert s.do_numeric_factorization(bad, raise_on_error=False).status == LinearSolverStatus.singular


def test_config5_shape_vs_oracle():
    """BASELINE config 5's block shape at full size -- 20 000 rows, 2 000 coupling columns (two of its 128 blocks):
    dense root fronts of ~4 082 rows through the cluster panel kernel and the DMMA update.  Against the closed-form
    inertia, the residual bar and the reference algorithm (oracle) on the same system."""
    m = EstimationModel(2, 2000, 4, 2000)
    assert m.block_dim == 20000
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s, x = _solve(kkt, rhs)
    st = s.backend.plan_stats(0)
    assert st["supernodes"] > 0 and not st["fell_back_dense"]
    assert s.get_inertia() == m.expected_inertia()
    assert _rel_residual(kkt, x, rhs) <= 1e-10
    o = SchurOracle()
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x_ref = o.solve(rhs).flatten()
    assert np.linalg.norm(x.flatten() - x_ref) / np.linalg.norm(x_ref) <= 1e-8
    assert abs(m.check_result(x) - m.check_result(o.solve(rhs))) <= 1e-8


def _lapack_inertia(dense):
    """Inertia from the pivots of LAPACK dsytrf (blocked Bunch-Kaufman)."""
    from scipy.linalg import lapack
    n = dense.shape[0]
    lw, _ = lapack.dsytrf_lwork(n, lower=1)
    ldu, ipiv, info = lapack.dsytrf(np.asfortranarray(dense), lower=1, lwork=int(lw), overwrite_a=1)
    assert info == 0
    pos = neg = zero = 0
    k = 0
    while k < n:
        if ipiv[k] < 0:   # 2x2 pivot (LAPACK convention, lower: ipiv[k] = ipiv[k+1] < 0)
            ev = np.linalg.eigvalsh(np.array([[ldu[k, k], ldu[k + 1, k]], [ldu[k + 1, k], ldu[k + 1, k + 1]]]))
            k += 2
        else:
            ev = np.array([ldu[k, k]])
            k += 1
        pos += int((ev > 0).sum()); neg += int((ev < 0).sum()); zero += int((ev == 0).sum())
    return pos, neg, zero


def _family_p_case(nb, scale, seed=7):
    from tests.helpers import stochastic_ipm_system
    n_x, n_eq, n_in, n_fs = int(10000 * scale), int(8000 * scale), int(1000 * scale), int(200 * scale)
    kkt, sizes = stochastic_ipm_system(seed, nb, n_x, n_eq, n_in, n_fs, same_pattern=True)
    rng = np.random.default_rng(0)
    return kkt, sizes, block_vector(rng.standard_normal(sum(sizes)), sizes)


def _check_blocks_against_superlu(kkt, rhs, x, nb):
    """Per block: residual <= max(1e-10, 2 x the reference SuperLU leaf's residual on the same block) and the two
    solutions agree to 1e-8 or to the forward-error bound cond_1(K) * eps.  Returns (leaves, worst leaf residual)."""
    import scipy.sparse.linalg as spl
    xc = np.asarray(x.get_block(nb))
    bnorm = np.linalg.norm(rhs.flatten())
    eps = np.finfo(float).eps
    leaves, worst_lu = [], 0.0
    for i in range(nb):
        K = kkt.get_block(i, i).tocsc()
        A = kkt.get_block(nb, i).tocsr()
        xi = np.asarray(x.get_block(i))
        bi = np.asarray(rhs.get_block(i)) - A.T @ xc
        lu = spl.splu(K)                                # the reference's leaf (scipy_interface.py:29,52) on this block
        y = lu.solve(bi)
        r_gpu = np.linalg.norm(K @ xi - bi) / bnorm
        r_lu = np.linalg.norm(K @ y - bi) / bnorm
        worst_lu = max(worst_lu, r_lu)
        assert r_gpu <= max(1e-10, 2.0 * r_lu), (i, r_gpu, r_lu)
        inv_norm = spl.onenormest(spl.LinearOperator(K.shape, matvec=lu.solve, rmatvec=lambda v: lu.solve(v, "T")))
        cond = inv_norm * spl.norm(K, 1)
        diff = np.linalg.norm(xi - y) / np.linalg.norm(y)
        assert diff <= max(1e-8, 10.0 * cond * eps), (i, diff, cond)
        leaves.append((K, A, lu))
    return leaves, worst_lu


def test_config4_shape_blocks_vs_superlu_leaf():
    """BASELINE config 4's block shape at full size: family-P scenarios of 20 200 rows (n_x 10 000, n_eq 8 000,
    n_in 1 000, 200 first-stage variables; two of the 1 024 scenarios).  Barrier diagonals over eight decades make
    some of these blocks numerically singular (|x_i| ~ 1e10 for |b| ~ 1e2), where no FP64 solver reaches 1e-10: the
    bar is "no worse than twice the reference's SuperLU leaf on the same block" (see the helper above).  The inertia
    is cross-checked at this size against the same library's dense Bunch-Kaufman path (a different pivoting
    algorithm on 20 400-row dense fronts; by Sylvester's law both must count the same signs) and, at a size LAPACK
    can factor, against dsytrf in test_config4_shape_inertia_vs_lapack."""
    nb = 2
    kkt, sizes, rhs = _family_p_case(nb, 1.0)
    assert sizes[0] == 20200
    s, x = _solve(kkt, rhs)
    st = s.backend.plan_stats(0)
    assert st["supernodes"] > 0 and not st["fell_back_dense"]
    _, worst_lu = _check_blocks_against_superlu(kkt, rhs, x, nb)
    assert _rel_residual(kkt, x, rhs) <= max(1e-10, 2.0 * worst_lu)
    inertia = s.get_inertia()
    del s
    d = B200SchurComplementLinearSolver(options={"sparse": 0})
    assert d.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful
    assert d.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    assert d.backend.plan_stats(0)["supernodes"] == 0
    assert d.get_inertia() == inertia
    assert inertia[2] == 0 and sum(inertia) == sum(sizes)


def test_config4_shape_inertia_vs_lapack():
    """Family-P scenarios at 0.35 x the config-4 block size (7 070 rows: the multifrontal path with delayed pivots):
    inertia = sum of LAPACK dsytrf pivot signs per block + the eigenvalue signs of S formed with the reference's
    leaf (Haynsworth additivity, explicit_schur_complement.py:157-172)."""
    nb = 2
    kkt, sizes, rhs = _family_p_case(nb, 0.35)
    s, x = _solve(kkt, rhs)
    assert s.backend.plan_stats(0)["supernodes"] > 0 and not s.backend.plan_stats(0)["fell_back_dense"]
    leaves, worst_lu = _check_blocks_against_superlu(kkt, rhs, x, nb)
    assert _rel_residual(kkt, x, rhs) <= max(1e-10, 2.0 * worst_lu)
    m_c = sizes[-1]
    S = np.zeros((m_c, m_c))
    tot = np.zeros(3, dtype=np.int64)
    for K, A, lu in leaves:
        tot += np.asarray(_lapack_inertia(K.toarray()), dtype=np.int64)
        S -= A @ lu.solve(A.T.toarray())
    tot += np.asarray(dense_inertia(0.5 * (S + S.T), "eigvalsh"), dtype=np.int64)
    assert s.get_inertia() == tuple(int(v) for v in tot)

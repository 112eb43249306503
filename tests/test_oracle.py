"""CPU: pin the oracle restatement against the reference's goldens and fixtures."""
import numpy as np
import pytest

from oracle import reference_loader
from oracle.kkt_generator import EstimationModel
from oracle.schur_oracle import (SchurOracle, LeafLU, dense_inertia, full_space_solve, solve_partitioned,
                                 sym_full)
from tests.helpers import block_vector, bordered_from_dense


@pytest.mark.parametrize("name", ["orig", "sym", "sym_q", "orig_q"])
def test_known_answer_8x8(known_answers, name):
    """reference test_explicit_schur_complement.py:13-55 (x vs dense solve, inertia vs eigenvalue signs)."""
    dense = known_answers[f"kat_{name}_dense"]
    rhs = known_answers[f"kat_{name}_rhs"]
    kkt = bordered_from_dense(dense, [2, 2, 2, 2])
    o = SchurOracle(compute_inertia=True)
    assert o.symbolic(kkt) == 0
    assert o.numeric(kkt) == 0
    x = o.solve(block_vector(rhs, [2, 2, 2, 2])).flatten()
    assert np.array_equal(x, known_answers[f"kat_{name}_x"])
    assert o.inertia() == tuple(known_answers[f"kat_{name}_inertia"])
    assert np.allclose(x, np.linalg.solve(dense, rhs))
    if name == "orig":
        assert np.allclose(x, [1.5, -.5, -.5, -.5, -.5, 1.5, -.5, -.5])
        assert o.inertia() == (6, 2, 0)


def test_known_answer_partitioned(known_answers):
    """reference test_mpi_explicit_schur_complement.py:19-115 on 1, 2, 3 emulated ranks."""
    dense = known_answers["kat_sym_q_dense"]
    rhs = known_answers["kat_sym_q_rhs"]
    kkt = bordered_from_dense(dense, [2, 2, 2, 2])
    for size in (1, 2, 3):
        st, x, inertia = solve_partitioned(kkt, block_vector(rhs, [2, 2, 2, 2]), size, compute_inertia=True)
        assert st == 0
        assert np.allclose(x.flatten(), known_answers["kat_mpi_x"], rtol=0, atol=1e-14)
        assert inertia == tuple(known_answers["kat_mpi_inertia"])
    assert np.array_equal(known_answers["kat_mpi_x"], known_answers["kat_mpi_x_refactor"])


def test_leaf_3x3(known_answers):
    """reference test_linear_solvers.py:63-79."""
    import scipy.sparse as sp
    leaf = LeafLU(compute_inertia=True)
    assert leaf.factor(sp.coo_matrix(known_answers["leaf_dense"])) == 0
    for r, x in zip(known_answers["leaf_rhs"], known_answers["leaf_x"]):
        assert np.array_equal(leaf.solve(r), x)
    assert leaf.inertia == tuple(known_answers["leaf_inertia"])
    for method in ("eigvalsh", "ldl"):
        assert dense_inertia(known_answers["leaf_dense"], method) == tuple(known_answers["leaf_inertia"])


def test_leaf_singular_status():
    import scipy.sparse as sp
    leaf = LeafLU()
    assert leaf.factor(sp.coo_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))) == 2


@pytest.mark.parametrize("tag", ["g_3_20_2_5", "g_4_60_3_10"])
def test_generator_small(generator_golden, tag):
    g = generator_golden
    args = tuple(int(v) for v in g[f"{tag}_args"])
    m = EstimationModel(*args)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    assert np.array_equal(np.tril(sym_full(kkt).toarray()), g[f"{tag}_kkt_lower"])
    assert np.array_equal(rhs.flatten(), g[f"{tag}_rhs"])
    o = SchurOracle(compute_inertia=True)
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x = o.solve(rhs)
    assert np.array_equal(x.flatten(), g[f"{tag}_x"])
    assert o.inertia() == tuple(g[f"{tag}_inertia"]) == m.expected_inertia()
    assert m.check_result(x) == g[f"{tag}_max_err"]
    for method in ("eigvalsh", "ldl"):
        assert dense_inertia(sym_full(kkt).toarray(), method) == m.expected_inertia()
    assert np.allclose(full_space_solve(kkt, rhs), x.flatten(), rtol=1e-9, atol=1e-9)


def test_generator_reference_golden(generator_golden):
    """reference examples/tests/test_examples.py:76-99: max_err == 0.3163456780448639 (7 places)."""
    g = generator_golden
    m = EstimationModel(3, 500, 12, 10)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    o = SchurOracle()
    o.symbolic(kkt)
    assert o.numeric(kkt) == 0
    x = o.solve(rhs)
    err = m.check_result(x)
    assert err == g["g_3_500_12_10_max_err"]
    assert abs(err - 0.3163456780448639) < 5e-8
    assert np.array_equal(np.asarray(x.get_block(3)), g["g_3_500_12_10_xc"])
    st, xp, _ = solve_partitioned(kkt, rhs, 2)
    assert st == 0 and abs(m.check_result(xp) - 0.3163456780448639) < 5e-8


def test_generator_config2_fixture(generator_golden):
    """BASELINE config 2 inputs (64 x 2000 x 50) are the reference generator's, bit for bit."""
    g = generator_golden
    m = EstimationModel(64, 150, 6, 50)
    assert m.block_dim == 2000
    k0 = m.blocks[0].kkt().tocsr()
    assert k0.nnz == g["g_64_150_6_50_k0_nnz"] == 8176
    assert np.abs(k0.data).sum() == g["g_64_150_6_50_k0_data_sum"]
    assert m.build_rhs().flatten().sum() == g["g_64_150_6_50_rhs_sum"]


@pytest.mark.reference
@pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not present (GPU box)")
def test_restatement_equals_reference_live():
    ref = reference_loader.load()
    for args in [(2, 15, 2, 4), (5, 30, 3, 7)]:
        m_ref, m = ref.Model(*args, 3), EstimationModel(*args)
        k_ref, k = m_ref.build_kkt(), m.build_kkt()
        assert (k_ref.tocsr() != k.tocsr()).nnz == 0
        s = ref.SchurComplementLinearSolver({i: ref.ScipyInterface(compute_inertia=True) for i in range(args[0])},
                                            ref.ScipyInterface(compute_inertia=True))
        s.do_symbolic_factorization(k_ref)
        s.do_numeric_factorization(k_ref)
        x_ref = s.do_back_solve(m_ref.build_rhs())
        o = SchurOracle(compute_inertia=True)
        o.symbolic(k)
        o.numeric(k)
        assert np.array_equal(o.solve(m.build_rhs()).flatten(), x_ref.flatten())
        assert o.inertia() == tuple(int(v) for v in s.get_inertia())


def test_pivot_sign_inertia_matches_scipy_ldl():
    """``dense_inertia(..., "ldl")`` calls LAPACK ``dsytrf`` directly; the count of pivot signs must be the one
    ``scipy.linalg.ldl`` (same routine, D rebuilt in Python) gives, on KKT-shaped matrices with zero (2,2) blocks,
    diagonals over eight decades and a structurally zero row -- and only the lower triangle may be read."""
    import scipy.linalg as sla

    def by_scipy(dense):
        n = dense.shape[0]
        _, d, _ = sla.ldl(np.tril(dense) + np.tril(dense, -1).T, lower=True)
        pos = neg = zero = k = 0
        while k < n:
            if k + 1 < n and d[k + 1, k] != 0.0:
                ev = np.linalg.eigvalsh(d[k:k + 2, k:k + 2])
                pos, neg, zero = pos + int((ev > 0).sum()), neg + int((ev < 0).sum()), zero + int((ev == 0).sum())
                k += 2
            else:
                pos, neg, zero = pos + int(d[k, k] > 0), neg + int(d[k, k] < 0), zero + int(d[k, k] == 0)
                k += 1
        return pos, neg, zero

    rng = np.random.default_rng(7)
    for t in range(120):
        n = int(rng.integers(1, 70))
        m = int(rng.integers(0, n + 1))
        H = rng.standard_normal((n, n))
        H = H + H.T if t % 3 else np.diag(10.0 ** rng.uniform(-4, 4, n))
        A = rng.standard_normal((m, n))
        K = np.block([[H, A.T], [A, np.zeros((m, m))]])
        if t % 7 == 0:
            K[:, -1] = 0
            K[-1, :] = 0
        lower_only = np.tril(K) + np.triu(rng.standard_normal(K.shape), 1)
        got = dense_inertia(lower_only, "ldl")
        assert got == by_scipy(K), t
        assert sum(got) == n + m

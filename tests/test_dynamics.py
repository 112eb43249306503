"""The dynamic (time-decomposed) structure: parapint's dynamics example end to end.

reference goldens: examples/tests/test_examples.py:35-57 -- nine optimal controls p(t), 7 places, from
``examples.dynamics.main`` with three time blocks on three MPI ranks and SciPy leaves.  The example is restated
with closed-form data in oracle/ipm.py (``dynamics_time_blocks`` + ``DynamicInterface``, following
examples/dynamics.py:37-100 and sc_ip_interface.py:22-1030); what reaches the linear solver is the reference's
nested layout: Q != 0, a 2x2 nested border block (N, i), backward-link multipliers inside the diagonal blocks.
"""
import numpy as np
import pytest

from oracle.ipm import DYNAMICS_GOLDEN_P, DynamicInterface, dynamics_time_blocks, ip_solve
from oracle.schur_oracle import OraclePlugin, sym_full


def _run(solver, **kw):
    blocks, starts, ends, p_times = dynamics_time_blocks(**kw)
    itf = DynamicInterface(blocks, starts, ends)
    out = ip_solve(itf, solver)
    p = {}
    for b, times in enumerate(p_times):
        n_x = itf.sc[b].nlp.n - len(times)
        for k, t in enumerate(times):
            p[t] = itf.sc[b].nlp.x[n_x + k]
    return itf, out, p


def test_dynamics_structure_is_the_reference_layout():
    blocks, starts, ends, _ = dynamics_time_blocks()
    itf = DynamicInterface(blocks, starts, ends)
    itf.set_barrier_parameter(0.1)
    for s in itf.sc:
        s.nlp.x = np.full(s.nlp.n, 0.5)
    kkt = itf.evaluate_primal_dual_kkt_matrix()
    rhs = itf.evaluate_primal_dual_kkt_rhs()
    N, n_s = 3, 1
    assert kkt.bshape == (N + 1, N + 1)
    # coupling part = forward-link multipliers of blocks 0..N-2, then the coupling variables (sc_ip_interface.py:335-357)
    Q = kkt.get_block(N, N)
    assert Q.bshape == (2, 2) and Q.shape == (2 * n_s * (N - 1), 2 * n_s * (N - 1))
    assert np.count_nonzero(Q.toarray()) == 2 * n_s * (N - 1)          # Q != 0: -Cf and its transpose
    for i in range(N):
        n = blocks[i].n + blocks[i].n_eq
        nb = 0 if i == 0 else n_s
        assert kkt.get_block(i, i).shape == (n + nb, n + nb)            # backward multipliers live in the block
        border = kkt.get_block(N, i)
        assert border.bshape == (2, 2) and border.shape == (2 * n_s * (N - 1), n + nb)
        assert rhs.get_block(i).nblocks == 2 and rhs.get_block(i).get_block(0).nblocks == 4
    assert rhs.get_block(N).nblocks == 2 and rhs.get_block(N).get_block(0).nblocks == N
    full = sym_full(kkt).toarray()
    assert np.allclose(full, full.T)


def test_dynamics_reference_algorithm_reaches_goldens():
    itf, out, p = _run(OraclePlugin())
    assert out["status"] == "optimal" and out["iterations"] == 10
    assert sorted(p) == sorted(DYNAMICS_GOLDEN_P)
    for t, gold in DYNAMICS_GOLDEN_P.items():
        assert abs(p[t] - gold) < 5e-8, (t, p[t], gold)                # assertAlmostEqual: 7 places
    # no regularisation is ever needed on this convex QP; inertia = (primal + coupling, multipliers, 0)
    assert all(r[1] == 0 for r in out["reg"])
    assert out["reg"][0][3] == itf.n_eq_constraints() + itf.n_ineq_constraints()
    # both inertia conventions agree here (no eigenvalue near the 1e-8 threshold)
    _, piv, p2 = _run(OraclePlugin(inertia_method="ldl"))
    assert piv["iterations"] == 10 and max(abs(p2[t] - p[t]) for t in p) < 1e-12


@pytest.mark.gpu
def test_dynamics_same_trajectory_on_b200():
    from parapint_b200 import B200SchurComplementLinearSolver
    _, ref, p_ref = _run(OraclePlugin())
    solver = B200SchurComplementLinearSolver()
    itf, out, p = _run(solver)
    assert out["status"] == "optimal"
    assert out["iterations"] == ref["iterations"] == 10
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    assert [(r[1], r[2], r[3], r[4]) for r in out["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]
    for t, gold in DYNAMICS_GOLDEN_P.items():
        assert abs(p[t] - gold) < 5e-8 and abs(p[t] - p_ref[t]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("num_time_blocks,nfe", [(6, 120), (15, 150)])
def test_dynamics_more_time_blocks_on_b200(num_time_blocks, nfe):
    """Longer chains of time blocks (the coupling system grows with the block count, config 3 in miniature)."""
    from parapint_b200 import B200SchurComplementLinearSolver
    kw = dict(num_finite_elements=nfe, num_time_blocks=num_time_blocks)
    _, ref, p_ref = _run(OraclePlugin(), **kw)
    _, out, p = _run(B200SchurComplementLinearSolver(), **kw)
    assert ref["status"] == out["status"] == "optimal"
    assert out["iterations"] == ref["iterations"]
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    assert max(abs(p[t] - p_ref[t]) for t in p_ref) < 1e-8

"""End-to-end: the plugin inside the (restated) interior-point loop on parapint's farmer example.

reference golden: examples/tests/test_examples.py:23-33 -- WHEAT 170, CORN 80, SUGAR_BEETS 250 (5 places)."""
import numpy as np
import pytest

from oracle.ipm import StochasticInterface, farmer_scenarios, ip_solve
from oracle.schur_oracle import OraclePlugin


# Inertia convention.  The reference's leaves disagree among themselves on "zero" eigenvalues (SURVEY.md 7,
# hard part 2): the SciPy leaf thresholds dense eigenvalues at 1e-8 (scipy_interface.py:42-44), MA27 / MUMPS
# report pivot signs (ma27_interface.py:201-203).  The B200 solver follows the pivot-sign convention, so the
# trajectory is compared with the reference algorithm under that convention (LAPACK dsytrf pivots); with the
# eigenvalue threshold the farmer run regularises on its first iteration and takes 52 instead of 53 iterations,
# reaching the same optimum.
def _run(solver):
    scen, first_stage, _ = farmer_scenarios()
    itf = StochasticInterface(scen, first_stage)
    out = ip_solve(itf, solver)
    return itf, out


def test_farmer_reference_algorithm_reaches_golden():
    itf, out = _run(OraclePlugin())
    assert out["status"] == "optimal"
    # devoted_acreage order in the model: WHEAT, CORN, SUGAR_BEETS
    for s in itf.sc:
        assert np.allclose(s.nlp.x[:3], [170.0, 80.0, 250.0], atol=5e-6)
    assert np.allclose(out["primals"][-3:], [170.0, 80.0, 250.0], atol=5e-6)
    assert abs(out["objective"] - (-108390.0)) < 0.2
    # under the SciPy leaf's eigenvalue threshold the inertia-correction loop is exercised on iteration 0
    assert any(r[1] > 0 for r in out["reg"]) and out["iterations"] == 52
    _, piv = _run(OraclePlugin(inertia_method="ldl"))
    assert piv["status"] == "optimal" and piv["iterations"] == 53
    assert abs(piv["objective"] - out["objective"]) < 1e-6 * abs(out["objective"])
    assert itf.n_eq_constraints() + itf.n_ineq_constraints() == 3 * (3 + 10)


@pytest.mark.gpu
def test_farmer_same_trajectory_on_b200():
    """north_star: identical IPM iteration counts and objective to 1e-8 against the reference algorithm."""
    from parapint_b200 import B200SchurComplementLinearSolver
    _, ref = _run(OraclePlugin(inertia_method="ldl"))
    itf, out = _run(B200SchurComplementLinearSolver())
    assert out["status"] == "optimal"
    assert out["iterations"] == ref["iterations"]
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * abs(ref["objective"])
    assert np.allclose(out["primals"][-3:], [170.0, 80.0, 250.0], atol=5e-6)
    # same regularisation decisions (integer inertia test of interior_point.py:379) at every iteration
    assert [(r[1], r[2], r[3], r[4]) for r in out["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]
    for (i1, o1, *_), (i2, o2, *_) in zip(out["history"], ref["history"]):
        assert i1 == i2 and abs(o1 - o2) <= 1e-6 * max(1.0, abs(o2))


@pytest.mark.gpu
@pytest.mark.parametrize("args", [(0, 4, 60, 20, 8, 5, 0.0), (1, 4, 60, 20, 8, 5, 0.3), (2, 6, 220, 120, 12, 6, 0.15)])
def test_stochastic_qp_same_trajectory_on_b200(args):
    """Scaled-up two-stage QPs; with negative curvature the inertia-correction loop refactorises dozens of times
    (interior_point.py:369-395) and the Hessian regularisation changes the COO pattern (interface.py:610-619), which
    the solver must notice.  The 370-row blocks of the last case take the multifrontal path."""
    from oracle.ipm import random_stochastic_qp
    from parapint_b200 import B200SchurComplementLinearSolver

    def run(solver):
        scen, first_stage = random_stochastic_qp(*args)
        return ip_solve(StochasticInterface(scen, first_stage), solver)

    ref = run(OraclePlugin(inertia_method="ldl"))
    solver = B200SchurComplementLinearSolver()
    out = run(solver)
    assert ref["status"] == out["status"] == "optimal"
    assert out["iterations"] == ref["iterations"]
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    assert [(r[1], r[2], r[3], r[4]) for r in out["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]
    if args[2] >= 200:
        assert solver.backend.plan_stats(0)["supernodes"] > 0

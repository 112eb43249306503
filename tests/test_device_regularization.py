"""Device-side inertia correction: the regularised KKT matrix travels as base + three shifts.

The inertia-correction loop of interior_point.py:369-395 must take the same decisions as with the interface's own
regularize_* methods (interface.py:590-619, sc_ip_interface.py:903-933,1736-1757), while the retries upload nothing
and never repeat the symbolic phase."""
import numpy as np
import pytest

from oracle.ipm import (DynamicInterface, StochasticInterface, device_regularized, dynamics_time_blocks, ip_solve,
                        random_stochastic_qp)
from oracle.schur_oracle import OraclePlugin, sym_full
from parapint_b200 import B200SchurComplementLinearSolver
from parapint_b200.regularization import RegularizedKKT
from tests.fake_backend import FakeBackend

ARGS = (1, 4, 40, 14, 6, 4, 0.3)    # nonconvex two-stage QP: dozens of retries


def _setup(cls, device):
    scen, fs = random_stochastic_qp(*ARGS)
    return (device_regularized(cls) if device else cls)(scen, fs)


def test_wrapper_equals_the_interface_regularisation():
    """materialize() of the wrapper == what the reference-style regularize_* produce, including the accumulation of the
    Hessian shift over retries and the SET semantics of the other two classes."""
    plain, dev = _setup(StochasticInterface, False), _setup(StochasticInterface, True)
    for itf in (plain, dev):
        itf.set_barrier_parameter(0.1)
        for s in itf.sc:
            s.nlp.x = np.full(s.nlp.n, 0.3)
    k0, k1 = plain.evaluate_primal_dual_kkt_matrix(), dev.evaluate_primal_dual_kkt_matrix()
    a, b = k0.copy(), k1.copy()
    for delta in (1e-8, 1e-7, 1e-6):
        a = plain.regularize_equality_gradient(kkt=a, coef=-delta, copy_kkt=False)
        a = plain.regularize_hessian(kkt=a, coef=delta, copy_kkt=False)
        b = dev.regularize_equality_gradient(kkt=b, coef=-delta, copy_kkt=False)
        b = dev.regularize_hessian(kkt=b, coef=delta, copy_kkt=False)
        assert isinstance(b, RegularizedKKT) and b.bshape == a.bshape
        assert np.allclose(sym_full(b.materialize()).toarray(), sym_full(a).toarray(), rtol=1e-14, atol=1e-22)
    assert b.shifts == (1e-8 + 1e-7 + 1e-6, -1e-6, 1e-6)
    # the dynamic layout: forward multipliers (coupling side) are class 2, the coupling variables class 3
    blocks, st, en, _ = dynamics_time_blocks()
    p, d = DynamicInterface(blocks, st, en), device_regularized(DynamicInterface)(blocks, st, en)
    for itf in (p, d):
        itf.set_barrier_parameter(0.1)
    a, b = p.evaluate_primal_dual_kkt_matrix().copy(), d.evaluate_primal_dual_kkt_matrix().copy()
    for delta in (1e-4, 1e-3):
        a = p.regularize_hessian(p.regularize_equality_gradient(a, -delta, False), delta, False)
        b = d.regularize_hessian(d.regularize_equality_gradient(b, -delta, False), delta, False)
    assert np.allclose(sym_full(b.materialize()).toarray(), sym_full(a).toarray(), rtol=1e-14, atol=1e-22)


def test_ipm_trajectory_and_counts_on_the_cpu_stand_in():
    """Host logic on CPU (numpy stand-in of the C ABI): same regularisation log and iterates as the plain interface;
    one value upload per IPM iteration instead of one per factorisation; one symbolic phase in total."""
    ref = ip_solve(_setup(StochasticInterface, False), OraclePlugin(inertia_method="eigvalsh"))
    itf = _setup(StochasticInterface, True)
    solver = B200SchurComplementLinearSolver(backend=FakeBackend(), regularization_classes=itf.regularization_classes())
    out = ip_solve(itf, solver)
    assert ref["status"] == out["status"] == "optimal" and out["iterations"] == ref["iterations"]
    assert [(r[1], r[2], r[3], r[4]) for r in out["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]
    assert abs(out["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    n_fact = len(out["reg"])
    assert n_fact > out["iterations"] + 10                       # the loop did retry
    assert solver.backend.value_uploads() == out["iterations"]   # one upload per KKT evaluation, none for retries
    assert solver.symbolic_calls == 1
    # classes learnt lazily from the first wrapper: one more symbolic phase, same answers
    itf2 = _setup(StochasticInterface, True)
    lazy = B200SchurComplementLinearSolver(backend=FakeBackend())
    out2 = ip_solve(itf2, lazy)
    assert out2["iterations"] == ref["iterations"] and lazy.symbolic_calls == 2
    assert [(r[1], r[2], r[3], r[4]) for r in out2["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]


@pytest.mark.gpu
@pytest.mark.parametrize("args", [ARGS, (2, 6, 220, 120, 12, 6, 0.15)])
def test_ipm_retries_are_refactor_only_on_b200(args):
    """zero re-symbolic calls and zero value uploads for the retries, same trajectory."""
    def run(solver, device):
        scen, fs = random_stochastic_qp(*args)
        itf = (device_regularized(StochasticInterface) if device else StochasticInterface)(scen, fs)
        return itf, ip_solve(itf, solver)

    _, ref = run(OraclePlugin(inertia_method="ldl"), False)
    plain_solver = B200SchurComplementLinearSolver()
    _, plain = run(plain_solver, False)
    scen, fs = random_stochastic_qp(*args)
    itf = device_regularized(StochasticInterface)(scen, fs)
    solver = B200SchurComplementLinearSolver(regularization_classes=itf.regularization_classes())
    out = ip_solve(itf, solver)
    for res in (plain, out):
        assert res["status"] == ref["status"] == "optimal" and res["iterations"] == ref["iterations"]
        assert [(r[1], r[2], r[3], r[4]) for r in res["reg"]] == [(r[1], r[2], r[3], r[4]) for r in ref["reg"]]
        assert abs(res["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))
    assert len(out["reg"]) > out["iterations"]                    # retries happened
    assert solver.symbolic_calls == 1                             # never re-analysed
    assert solver.backend.value_uploads() == out["iterations"]    # uploads: one per KKT evaluation
    assert plain_solver.backend.value_uploads() == len(plain["reg"])   # the plain interface uploads for every retry
    assert plain_solver.symbolic_calls > 1                        # ... and re-analyses when the Hessian pattern grows


@pytest.mark.gpu
def test_dynamic_layout_shifts_on_b200():
    """The dynamic layout's classes (forward multipliers on the coupling side are class 2, interface states class 3):
    the device applies the three shifts exactly where the interface's regularize_* put them -- same solution and
    inertia as the reference algorithm on the materialised matrix, with and without the sparse coupling path; a second
    factorisation with other shifts re-uses the values on the device."""
    from oracle.schur_oracle import SchurOracle
    blocks, st, en, _ = dynamics_time_blocks(num_finite_elements=120, num_time_blocks=12)
    itf = device_regularized(DynamicInterface)(blocks, st, en)
    itf.set_barrier_parameter(0.1)
    rng = np.random.default_rng(0)
    for s_ in itf.sc:
        s_.nlp.x = rng.uniform(0.2, 1.5, s_.nlp.n)
    kkt = itf.evaluate_primal_dual_kkt_matrix()
    rhs = itf.evaluate_primal_dual_kkt_rhs()
    for options in ({}, {"coupling_min_sparse": 8}):
        solver = B200SchurComplementLinearSolver(regularization_classes=itf.regularization_classes(), options=options)
        assert solver.do_symbolic_factorization(kkt).status.value == 0
        assert solver.do_numeric_factorization(kkt).status.value == 0
        base_inertia = solver.get_inertia()
        reg = kkt.copy()
        for delta in (1e-3, 1e-1):
            reg = itf.regularize_equality_gradient(kkt=reg, coef=-delta, copy_kkt=False)
            reg = itf.regularize_hessian(kkt=reg, coef=delta, copy_kkt=False)
            assert solver.do_numeric_factorization(reg).status.value == 0
            x = solver.do_back_solve(rhs).flatten()
            full = reg.materialize()
            o = SchurOracle(compute_inertia=True, inertia_method="ldl")
            o.symbolic(full)
            assert o.numeric(full) == 0
            x_ref = o.solve(rhs).flatten()
            assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) <= 1e-8
            assert solver.get_inertia() == o.inertia()
        assert solver.symbolic_calls == 1 and solver.backend.value_uploads() == 1
        # back to the unregularised matrix: shifts are cleared, values still on the device
        assert solver.do_numeric_factorization(kkt).status.value == 0 and solver.get_inertia() == base_inertia
        assert solver.backend.value_uploads() == 1

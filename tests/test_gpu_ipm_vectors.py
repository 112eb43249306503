"""GPU parity of the interior-point vector kernels against the oracle's restatement of
``parapint/algorithms/interior_point.py:174-317,655-758`` and ``parapint/interfaces/interface.py:548-570``.

Bar: minima / maxima (step lengths, infeasibilities, complementarity) BIT-EXACT -- the kernels follow the reference's
NumPy expressions operation by operation and a min / max does not depend on the order of the reduction; the two
scaling factors, which contain sums, to 1e-13 relative."""
import numpy as np
import pytest

from oracle import ipm as O

pytestmark = pytest.mark.gpu


def _random_group(rng, n, p_inf=0.3):
    """Iterates strictly inside random bounds, some of them infinite, multipliers zero on infinite bounds
    (process_init_duals_*, interior_point.py:789-797), steps of either sign with exact zeros sprinkled in."""
    lb = rng.uniform(-5, 0, n)
    ub = lb + rng.uniform(0.5, 6, n)
    x = lb + (ub - lb) * rng.uniform(0.01, 0.99, n)
    lb[rng.random(n) < p_inf] = -np.inf
    ub[rng.random(n) < p_inf] = np.inf
    zl = np.where(np.isneginf(lb), 0.0, 10.0 ** rng.uniform(-8, 2, n))
    zu = np.where(np.isinf(ub), 0.0, 10.0 ** rng.uniform(-8, 2, n))
    dx = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 1, n)
    dx[rng.random(n) < 0.05] = 0.0
    return x, dx, lb, ub, zl, zu


def _ref_ftb(tau, barrier, x, dx, lb, ub, zl, zu):
    dzl = (barrier - zl * dx) / (x - lb) - zl          # interface.py:548-553
    dzu = (barrier + zu * dx) / (ub - x) - zu          # :555-559
    a_p = min(O._ftb_lb(tau, x, dx, lb), O._ftb_ub(tau, x, dx, ub))
    a_d = min(O._ftb_lb(tau, zl, dzl, np.zeros_like(zl)), O._ftb_lb(tau, zu, dzu, np.zeros_like(zu)))
    return a_p, a_d, dzl, dzu


@pytest.mark.parametrize("n", [0, 1, 2, 7, 1000, 100003, 1 << 21])
def test_kernels_vs_numpy(n):
    import torch
    from parapint_b200.ipm_vectors import IpmKernels
    k = IpmKernels()
    rng = np.random.default_rng(n)
    x, dx, lb, ub, zl, zu = _random_group(rng, n)
    tau, barrier = 0.995, 3.7e-3
    with np.errstate(all="ignore"):
        a_p, a_d, dzl, dzu = _ref_ftb(tau, barrier, x, dx, lb, ub, zl, zu)
    d = [k.to_device(v) for v in (x, dx, lb, ub, zl, zu)]
    out = torch.ones(2, dtype=torch.float64, device=k.device)
    k.fraction_to_boundary(out, tau, barrier, *d)
    got = out.cpu().numpy()
    assert got[0] == a_p and got[1] == a_d, (got, a_p, a_d)          # bit-exact

    # complementarity terms, interior_point.py:241-251 / :274-315
    out6 = torch.zeros(6, dtype=torch.float64, device=k.device)
    k.complementarity(out6, barrier, d[0], d[2], d[3], d[4], d[5])
    lb_m, ub_m = lb.copy(), ub.copy()
    lb_m[np.isneginf(lb)] = 0
    ub_m[np.isinf(ub)] = 0
    r_l = (x - lb_m) * zl - barrier
    r_u = (ub_m - x) * zu - barrier
    r_l[np.isneginf(lb)] = 0
    r_u[np.isinf(ub)] = 0
    o = out6.cpu().numpy()
    assert o[0] == O._max_abs(r_l) and o[1] == O._max_abs(r_u)      # bit-exact
    assert o[4] == np.isfinite(lb).sum() and o[5] == np.isfinite(ub).sum()
    assert np.isclose(o[2], np.abs(zl).sum(), rtol=1e-13, atol=0) and np.isclose(o[3], np.abs(zu).sum(), rtol=1e-13, atol=0)

    # max |a - b| and sum |a|
    out2 = torch.zeros(2, dtype=torch.float64, device=k.device)
    k.max_abs(out2, d[0], d[1])
    o = out2.cpu().numpy()
    assert o[0] == O._max_abs(x - dx) and np.isclose(o[1], np.abs(x).sum(), rtol=1e-13, atol=0)

    # step update, interior_point.py:588-595,619-626: every entry bit-exact
    alpha = torch.tensor([min(a_p, 1.0), min(a_d, 1.0), 1.0], dtype=torch.float64, device=k.device)
    k.step(alpha, barrier, d[0], d[1], d[2], d[3], d[4], d[5])
    ap, ad = min(a_p, 1.0), min(a_d, 1.0)
    with np.errstate(all="ignore"):
        assert np.array_equal(d[0].cpu().numpy(), x + 1.0 * (ap * dx))
        assert np.array_equal(d[4].cpu().numpy(), zl + 1.0 * (ad * dzl), equal_nan=True)
        assert np.array_equal(d[5].cpu().numpy(), zu + 1.0 * (ad * dzu), equal_nan=True)
    y, dy = rng.standard_normal(n), rng.standard_normal(n)
    yd = k.to_device(y)
    k.axpy(alpha, 1, yd, k.to_device(dy))
    assert np.array_equal(yd.cpu().numpy(), y + 1.0 * (ad * dy))


def test_unaligned_views_take_the_scalar_path():
    """Vectors that start at an odd element (8-byte aligned only) must give the same answers."""
    import torch
    from parapint_b200.ipm_vectors import IpmKernels
    k = IpmKernels()
    rng = np.random.default_rng(5)
    n = 4099
    vecs = _random_group(rng, n)
    base = [k.to_device(np.concatenate([[0.0], v])) for v in vecs]
    views = [b[1:] for b in base]
    assert views[0].data_ptr() % 16 == 8
    out_a, out_b = (torch.ones(2, dtype=torch.float64, device=k.device) for _ in range(2))
    k.fraction_to_boundary(out_a, 0.99, 1e-2, *[v.contiguous() for v in views])
    k.fraction_to_boundary(out_b, 0.99, 1e-2, *[k.to_device(v) for v in vecs])
    assert torch.equal(out_a, out_b)


def _interface_mid_solve(seed=3):
    """A stochastic QP interface a few iterations into ``ip_solve`` (iterates, multipliers and steps all set)."""
    from oracle.schur_oracle import OraclePlugin
    scen, fs = O.random_stochastic_qp(seed, 6, 40, 20, 8, 4, 0.1)
    itf = O.StochasticInterface(scen, fs)
    opts = O.IPOptions()
    opts.max_iter = 6
    O.ip_solve(itf, OraclePlugin(inertia_method="ldl"), opts)
    return itf


@pytest.mark.parametrize("expose_barrier", [True, False])
def test_drop_in_functions_vs_oracle(expose_barrier):
    """``fraction_to_the_boundary(interface, tau)`` and ``check_convergence(interface, barrier, error_scaling)`` with
    the reference's signatures, on an interface in the state the loop leaves it in."""
    from parapint_b200 import ipm_vectors as V
    itf = _interface_mid_solve()
    barrier = itf.sc[0].barrier
    if expose_barrier:
        itf._barrier = barrier          # as parapint's InteriorPointInterface keeps it (interface.py:365)
    ref = O.fraction_to_the_boundary(itf, 1 - barrier)
    got = V.fraction_to_the_boundary(itf, 1 - barrier)
    assert got == ref, (got, ref)
    for b in (0.0, barrier):
        ref = O.check_convergence(itf, b, 100.0)
        got = V.check_convergence(itf, b, 100.0)
        assert got[0] == ref[0]
        assert np.isclose(got[1], ref[1], rtol=1e-13, atol=0) and np.isclose(got[2], ref[2], rtol=1e-13, atol=0)


def test_device_resident_iterates_follow_the_loop():
    """DeviceIpmVectors: load once, then step lengths, convergence terms and the update on the device; the updated
    iterates equal the loop's own update (interior_point.py:619-626) bit for bit."""
    from parapint_b200.ipm_vectors import DeviceIpmVectors
    itf = _interface_mid_solve(seed=4)
    barrier = itf.sc[0].barrier
    dv = DeviceIpmVectors().load(itf).set_steps(itf)
    a_p, a_d = dv.fraction_to_the_boundary(1 - barrier, barrier)
    assert (a_p, a_d) == O.fraction_to_the_boundary(itf, 1 - barrier)
    ref = O.check_convergence(itf, barrier, 100.0)
    c_inf, dual_scaling, compl_scaling = dv.complementarity(barrier, 100.0)
    assert np.isclose(c_inf / compl_scaling, ref[2], rtol=1e-13, atol=0)
    expect = {"primals": itf.get_primals() + a_p * itf.get_delta_primals(),
              "slacks": itf.get_slacks() + a_p * itf.get_delta_slacks(),
              "duals_eq": itf.get_duals_eq() + a_d * itf.get_delta_duals_eq(),
              "duals_ineq": itf.get_duals_ineq() + a_d * itf.get_delta_duals_ineq(),
              "duals_primals_lb": itf.get_duals_primals_lb() + a_d * itf.get_delta_duals_primals_lb(),
              "duals_primals_ub": itf.get_duals_primals_ub() + a_d * itf.get_delta_duals_primals_ub(),
              "duals_slacks_lb": itf.get_duals_slacks_lb() + a_d * itf.get_delta_duals_slacks_lb(),
              "duals_slacks_ub": itf.get_duals_slacks_ub() + a_d * itf.get_delta_duals_slacks_ub()}
    dv.take_step(barrier)
    for name, ref_v in expect.items():
        assert np.array_equal(dv.download(name), ref_v), name


def test_ip_solve_trajectory_with_device_vector_kernels():
    """The whole loop with BOTH the B200 linear solver and the device vector kernels: iteration count, objective and
    the infeasibility history equal the all-host run (reference algorithm + oracle leaf)."""
    from oracle.schur_oracle import OraclePlugin
    from parapint_b200 import B200SchurComplementLinearSolver, ipm_vectors as V
    args = (7, 5, 30, 14, 6, 3, 0.0)
    scen, fs = O.random_stochastic_qp(*args)
    ref = O.ip_solve(O.StochasticInterface(scen, fs), OraclePlugin(inertia_method="ldl"))
    scen, fs = O.random_stochastic_qp(*args)
    got = O.ip_solve(O.StochasticInterface(scen, fs), B200SchurComplementLinearSolver(),
                     check_convergence=V.check_convergence, fraction_to_the_boundary=V.fraction_to_the_boundary)
    assert got["status"] == ref["status"] == "optimal"
    assert got["iterations"] == ref["iterations"]
    assert abs(got["objective"] - ref["objective"]) <= 1e-8 * max(1.0, abs(ref["objective"]))

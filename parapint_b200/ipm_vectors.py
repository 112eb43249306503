"""Interior-point vector kernels on the device.

Between two linear solves parapint's interior-point loop makes a dozen O(n) NumPy passes over its iterates:
``fraction_to_the_boundary`` (reference ``parapint/algorithms/interior_point.py:677-758`` with the helpers
``:655-674``), the bound-multiplier steps of the interface (``parapint/interfaces/interface.py:548-570``), the
complementarity / scaling terms of ``check_convergence`` (``:174-317``) and the step update (``:587-626``).  Here every
one of them is a single streaming pass over device-resident vectors (``csrc/ipmvec.cuh`` behind the ``pp_ipm_*`` entry
points of ``include/parapint_b200.h``), with the reduction finished on the device.

Two ways in:

* :class:`DeviceIpmVectors` keeps the iterates of an interface in HBM and offers the passes as methods -- the
  B200-first form: nothing crosses PCIe but two or three doubles per call.
* :func:`fraction_to_the_boundary` and :func:`check_convergence` have the signatures of the reference's functions of
  the same names and take any interface with the reference's getters (host vectors): drop-in, at the price of
  uploading the vectors of that call.

There is no CPU fallback: without the built library or a CUDA device these raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native


def _flat(v):
    """Host vector of an interface getter (ndarray or a PyNumero-style block vector) as a float64 array."""
    if hasattr(v, "flatten") and not isinstance(v, np.ndarray):
        v = v.flatten()
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())


class IpmKernels:
    """The ``pp_ipm_*`` entry points on torch device tensors of one CUDA device (one workspace, one stream order)."""

    def __init__(self, device=None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("the interior-point vector kernels need a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.lib = native.load()
        index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", index)
        nbytes = int(self.lib.pp_ipm_workspace_bytes())
        self.workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device=self.device)
        self.out = torch.zeros(16, dtype=torch.float64, device=self.device)

    # -- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _p(self, t):
        if t is None:
            return None
        if t.dtype != self.torch.float64 or not t.is_contiguous() or t.device != self.device:
            raise ValueError("expected a contiguous float64 tensor on " + str(self.device))
        return C.c_void_p(t.data_ptr())

    def _check(self, code, what):
        if code != 0:
            raise RuntimeError(f"{what} failed: {native.last_error()}")

    def to_device(self, v):
        return self.torch.from_numpy(_flat(v)).to(self.device)

    def fill(self, t, value):
        self._check(self.lib.pp_ipm_fill(self._p(t), t.numel(), float(value), self._stream()), "pp_ipm_fill")

    # -- the passes --------------------------------------------------------------------------
    def fraction_to_boundary(self, out2, tau, barrier, x, dx, lb, ub, zl, zu):
        """``out2[0] = min(out2[0], alpha_primal of this group)``, ``out2[1] = min(out2[1], alpha_dual)``."""
        self._check(self.lib.pp_ipm_fraction_to_boundary(x.numel(), float(tau), float(barrier), self._p(x), self._p(dx),
                                                         self._p(lb), self._p(ub), self._p(zl), self._p(zu), self._p(out2),
                                                         self._p(self.workspace), self._stream()),
                    "pp_ipm_fraction_to_boundary")

    def complementarity(self, out6, barrier, x, lb, ub, zl, zu):
        self._check(self.lib.pp_ipm_complementarity(x.numel(), float(barrier), self._p(x), self._p(lb), self._p(ub),
                                                    self._p(zl), self._p(zu), self._p(out6), self._p(self.workspace),
                                                    self._stream()), "pp_ipm_complementarity")

    def max_abs(self, out2, a, b=None):
        self._check(self.lib.pp_ipm_max_abs(a.numel(), self._p(a), self._p(b), self._p(out2), self._p(self.workspace),
                                            self._stream()), "pp_ipm_max_abs")

    def step(self, alpha3, barrier, x, dx, lb, ub, zl, zu):
        self._check(self.lib.pp_ipm_step(x.numel(), self._p(alpha3), float(barrier), self._p(x), self._p(dx), self._p(lb),
                                         self._p(ub), self._p(zl), self._p(zu), self._stream()), "pp_ipm_step")

    def axpy(self, alpha3, which, y, dy):
        self._check(self.lib.pp_ipm_axpy(y.numel(), self._p(alpha3), int(which), self._p(y), self._p(dy), self._stream()),
                    "pp_ipm_axpy")


class DeviceIpmVectors:
    """Iterates of an interior-point interface kept in HBM.

    ``load(interface)`` uploads the bounds (once) and the current iterates; ``set_steps`` uploads the primal / slack /
    multiplier steps of the last linear solve; after that ``fraction_to_the_boundary``, ``complementarity`` and
    ``take_step`` run entirely on the device.  Names follow the reference interface
    (``parapint/interfaces/interface.py:10-360``)."""

    GROUPS = (("primals", "primals_lb", "primals_ub", "duals_primals_lb", "duals_primals_ub", "delta_primals"),
              ("slacks", "ineq_lb", "ineq_ub", "duals_slacks_lb", "duals_slacks_ub", "delta_slacks"))

    def __init__(self, device=None, kernels=None, comm=None):
        """``comm``: a :class:`parapint_b200.comm.Communicator` when every rank holds the iterates of its own blocks
        (the reference's ``MPIBlockVector`` reductions, ``interior_point.py:253-315`` on distributed vectors): step
        lengths and maxima are then MIN / MAX-reduced, sums and counts SUM-reduced over the ranks -- two small
        all-reduces per call.  Replicated vectors (the coupling variables) must be loaded on one rank only."""
        self.k = kernels if kernels is not None else IpmKernels(device)
        self.comm = comm
        self.v = {}
        self.alpha = self.k.torch.ones(3, dtype=self.k.torch.float64, device=self.k.device)

    def _reduce(self, t, maxima, sums):
        """In-place reduction over the ranks: ``t[maxima]`` by MAX, ``t[sums]`` by SUM (slices of one device tensor)."""
        if self.comm is None or self.comm.size == 1:
            return
        if maxima is not None:
            self.comm.allreduce_max_(t[maxima])
        if sums is not None:
            self.comm.allreduce_sum_(t[sums])

    def load(self, interface):
        up = self.k.to_device
        for name in ("primals_lb", "primals_ub", "ineq_lb", "ineq_ub"):
            self.v[name] = up(getattr(interface, name)())
        for name in ("primals", "slacks", "duals_eq", "duals_ineq", "duals_primals_lb", "duals_primals_ub",
                     "duals_slacks_lb", "duals_slacks_ub"):
            self.v[name] = up(getattr(interface, "get_" + name)())
        return self

    def set_steps(self, interface):
        up = self.k.to_device
        for name in ("delta_primals", "delta_slacks", "delta_duals_eq", "delta_duals_ineq"):
            self.v[name] = up(getattr(interface, "get_" + name)())
        return self

    def fraction_to_the_boundary(self, tau, barrier):
        """(alpha_primal_max, alpha_dual_max) of ``interior_point.py:677-758``; they also stay in ``self.alpha[:2]``
        for :meth:`take_step`."""
        k, v = self.k, self.v
        k.fill(self.alpha, 1.0)
        for x, lb, ub, zl, zu, dx in self.GROUPS:
            k.fraction_to_boundary(self.alpha, tau, barrier, v[x], v[dx], v[lb], v[ub], v[zl], v[zu])
        if self.comm is not None and self.comm.size > 1:      # min over the ranks = -max(-x)
            self.alpha[:2].neg_()
            self._reduce(self.alpha, slice(0, 2), None)
            self.alpha[:2].neg_()
        a = self.alpha[:2].cpu().numpy()
        return float(a[0]), float(a[1])

    def complementarity(self, barrier, error_scaling):
        """(complimentarity_inf before scaling, dual_scaling, compl_scaling) of ``interior_point.py:241-251,274-315``."""
        k, v = self.k, self.v
        out = k.out
        k.fill(out, 0.0)
        for x, lb, ub, zl, zu, _ in self.GROUPS:
            k.complementarity(out[:6], barrier, v[x], v[lb], v[ub], v[zl], v[zu])
        k.max_abs(out[6:8], v["duals_eq"])
        k.max_abs(out[8:10], v["duals_ineq"])
        n_eq, n_in = v["duals_eq"].numel(), v["duals_ineq"].numel()
        if self.comm is not None and self.comm.size > 1:
            # regroup as [maxima | sums and counts] so that two all-reduces do: o[0], o[1] MAX; the rest SUM
            out[10], out[11] = float(n_eq), float(n_in)
            packed = k.torch.cat([out[0:2], out[2:6], out[7:8], out[9:10], out[10:12]])
            self._reduce(packed, slice(0, 2), slice(2, 10))
            p = packed.cpu().numpy()
            o = np.zeros(10)
            o[0:2], o[2:6], o[7], o[9] = p[0:2], p[2:6], p[6], p[7]
            n_eq, n_in = int(round(p[8])), int(round(p[9]))
        else:
            o = out[:10].cpu().numpy()
        return _scalings(o, n_eq, n_in, error_scaling)

    def take_step(self, barrier, line_search_step=1.0, unified_step=False):
        """``interior_point.py:571-574,587-626`` with the step lengths already in ``self.alpha`` (device)."""
        k, v = self.k, self.v
        if unified_step:
            self.alpha[:2] = self.alpha[:2].min()
        self.alpha[2] = float(line_search_step)
        for x, lb, ub, zl, zu, dx in self.GROUPS:
            k.step(self.alpha, barrier, v[x], v[dx], v[lb], v[ub], v[zl], v[zu])
        k.axpy(self.alpha, 1, v["duals_eq"], v["delta_duals_eq"])
        k.axpy(self.alpha, 1, v["duals_ineq"], v["delta_duals_ineq"])

    def download(self, name):
        return self.v[name].cpu().numpy()


def _scalings(o, n_eq, n_ineq, error_scaling):
    """Fold the device reductions as ``interior_point.py:274-315`` does (``o`` = the ten accumulated values)."""
    compl_inf = max(o[0], o[1])
    bound_duals = o[2] + o[3]
    finite = o[4] + o[5]
    with np.errstate(divide="ignore", invalid="ignore"):
        dual_scaling = np.float64(o[7] + o[9] + bound_duals) / np.float64(n_eq + n_ineq + finite)
        compl_scaling = np.float64(bound_duals) / np.float64(finite)
    dual_scaling = max(error_scaling, dual_scaling) / error_scaling
    compl_scaling = max(error_scaling, compl_scaling) / error_scaling
    return float(compl_inf), float(dual_scaling), float(compl_scaling)


_kernels = {}


def _default_kernels():
    import torch

    index = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if index not in _kernels:
        _kernels[index] = IpmKernels()
    return _kernels[index]


def fraction_to_the_boundary(interface, tau):
    """Drop-in for ``parapint.algorithms.interior_point.fraction_to_the_boundary`` (``:677-758``): same arguments, same
    return value ``(alpha_primal_max, alpha_dual_max)``.  The bound-multiplier steps are formed on the device from the
    barrier parameter, the multipliers and the primal / slack steps (``interface.py:548-570``); interfaces that do not
    expose their barrier parameter as ``_barrier`` / ``barrier`` are asked for ``get_delta_duals_*`` instead."""
    k = _default_kernels()
    barrier = getattr(interface, "_barrier", getattr(interface, "barrier", None))
    up = k.to_device
    out = k.out[:2]
    k.fill(out, 1.0)
    groups = (("primals", "primals_lb", "primals_ub", "duals_primals_lb", "duals_primals_ub", "delta_primals"),
              ("slacks", "ineq_lb", "ineq_ub", "duals_slacks_lb", "duals_slacks_ub", "delta_slacks"))
    for x, lb, ub, zl, zu, dx in groups:
        xv, dxv = up(getattr(interface, "get_" + x)()), up(getattr(interface, "get_" + dx)())
        lbv, ubv = up(getattr(interface, lb)()), up(getattr(interface, ub)())
        zlv, zuv = up(getattr(interface, "get_" + zl)()), up(getattr(interface, "get_" + zu)())
        if barrier is not None:
            k.fraction_to_boundary(out, tau, barrier, xv, dxv, lbv, ubv, zlv, zuv)
        else:
            # multiplier steps given by the interface: two passes of the same kernel, multipliers as the "primal"
            # vector against the bounds [0, inf) with their own steps (no multipliers of the multipliers: zeros)
            zero, inf = k.torch.zeros_like(xv), k.torch.full_like(xv, float("inf"))
            tmp = k.torch.ones(2, dtype=k.torch.float64, device=k.device)
            k.fraction_to_boundary(tmp, tau, 0.0, xv, dxv, lbv, ubv, zero, zero)
            out[0] = k.torch.minimum(out[0], tmp[0])
            for z, dz in ((zlv, "get_delta_" + zl), (zuv, "get_delta_" + zu)):
                k.fill(tmp, 1.0)
                k.fraction_to_boundary(tmp, tau, 0.0, z, up(getattr(interface, dz)()), zero, inf, zero, zero)
                out[1] = k.torch.minimum(out[1], tmp[0])
    a = out.cpu().numpy()
    return float(a[0]), float(a[1])


def check_convergence(interface, barrier, error_scaling, timer=None):
    """Drop-in for ``parapint.algorithms.interior_point.check_convergence`` (``:174-317``): same arguments, same return
    value ``(primal_inf, dual_inf / dual_scaling, complimentarity_inf / compl_scaling)``.  The interface evaluates its
    functions and derivatives (``:192-207``) and forms the gradient of the Lagrangian (``:231-241``: sparse
    Jacobian-transpose products on the host); every bound / multiplier pass and every reduction runs on the device."""
    k = _default_kernels()
    up = k.to_device
    slacks = interface.get_slacks()
    grad_obj = interface.get_obj_factor() * interface.evaluate_grad_objective()
    jac_eq, jac_ineq = interface.evaluate_jacobian_eq(), interface.evaluate_jacobian_ineq()
    duals_eq, duals_ineq = interface.get_duals_eq(), interface.get_duals_ineq()
    zl, zu = interface.get_duals_primals_lb(), interface.get_duals_primals_ub()
    sl, su = interface.get_duals_slacks_lb(), interface.get_duals_slacks_ub()
    grad_lag_primals = grad_obj + jac_eq.transpose() * duals_eq
    grad_lag_primals += jac_ineq.transpose() * duals_ineq
    grad_lag_primals -= zl
    grad_lag_primals += zu
    grad_lag_slacks = (-duals_ineq - sl + su)

    out = k.out
    k.fill(out, 0.0)
    d_eq, d_in = up(duals_eq), up(duals_ineq)
    k.complementarity(out[:6], barrier, up(interface.get_primals()), up(interface.primals_lb()), up(interface.primals_ub()),
                      up(zl), up(zu))
    s_dev = up(slacks)
    k.complementarity(out[:6], barrier, s_dev, up(interface.ineq_lb()), up(interface.ineq_ub()), up(sl), up(su))
    k.max_abs(out[6:8], d_eq)
    k.max_abs(out[8:10], d_in)
    k.max_abs(out[10:12], up(interface.evaluate_eq_constraints()))
    k.max_abs(out[10:12], up(interface.evaluate_ineq_constraints()), s_dev)      # max |g(x) - s|  (:204)
    k.max_abs(out[12:14], up(grad_lag_primals))
    k.max_abs(out[12:14], up(grad_lag_slacks))
    o = out[:14].cpu().numpy()
    compl_inf, dual_scaling, compl_scaling = _scalings(o, d_eq.numel(), d_in.numel(), error_scaling)
    return float(o[10]), float(o[12]) / dual_scaling, compl_inf / compl_scaling

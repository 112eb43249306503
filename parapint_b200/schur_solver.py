"""``B200SchurComplementLinearSolver`` -- drop-in for parapint's Schur-complement linear solvers.

Replaces ``SchurComplementLinearSolver``
(reference ``parapint/linalg/schur_complement/explicit_schur_complement.py:16``) and
``MPISchurComplementLinearSolver`` (``mpi_explicit_schur_complement.py:128``) *together with*
their per-block sub-solvers behind the same ``LinearSolverInterface``
(``base_linear_solver_interface.py:5-56``), so it plugs into ``ip_solve`` as
``options.linalg.solver`` (``algorithms/interior_point.py:347,371,387,544,566``).

Host code is Python; device buffers, streams and collectives come from PyTorch; all arithmetic
runs in the hand-written sm_100a kernels behind the C ABI (``include/parapint_b200.h``).  There is
no CPU fallback: constructing the solver without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import logging
from typing import Optional

import numpy as np

from . import native, structure
from .comm import Communicator
from .interface import LinearSolverInterface, LinearSolverResults, LinearSolverStatus
from .regularization import RegularizedKKT

_OK = (LinearSolverStatus.successful, LinearSolverStatus.warning)
# severity order of the status values for cross-rank agreement
_SEVERITY = {0: 0, 4: 1, 1: 2, 2: 3, 3: 4}
_FROM_SEVERITY = {v: k for k, v in _SEVERITY.items()}
SCHUR_TAIL = 8  # PP_SCHUR_TAIL: status / inertia words appended to the Schur buffer


class _NullTimer:
    def start(self, name):
        pass

    def stop(self, name):
        pass


class CudaBackend:
    """Owns the native handle and the PyTorch device / pinned buffers for one GPU."""

    def __init__(self, device: Optional[int] = None, options: Optional[dict] = None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("B200SchurComplementLinearSolver needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.lib = native.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.handle = C.c_void_p()
        if self._check(self.lib.pp_create(self.device_index, C.byref(self.handle)), "pp_create") != 0 or not self.handle:
            raise RuntimeError(f"pp_create failed: {native.last_error()}")
        for key, val in (options or {}).items():
            self._check(self.lib.pp_set_option(self.handle, key.encode(), float(val)), "pp_set_option")
        self.st = None
        self._auto_residual = False
        with torch.cuda.device(self.device):
            self.ints = torch.zeros(4, dtype=torch.int64, device=self.device)   # status agreement across ranks
        self.last_error = ""
        self.failed = False

    def _check(self, code, what):
        """Only API misuse (PP_MISUSE: wrong call order, null pointer -- the same mistake on every rank) raises
        here.  Every other code, PP_ERROR included, is a *status* of this rank that the solver first agrees
        across ranks (``_finish`` / the tail of the Schur all-reduce) and only then returns or raises on."""
        if code < 0:
            raise RuntimeError(f"{what}: {native.last_error()}")
        if code == 3:
            self.last_error = f"{what} failed: {native.last_error()}"
            self.failed = True
        return code

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _release_exchange(self):
        comm = getattr(self, "_comm", None)
        if comm is not None:
            for bufs in (getattr(self, "_schur_bufs", None), getattr(self, "_rc_bufs", None), getattr(self, "_res_bufs", None)):
                comm.release_buffers(bufs)
        self._schur_bufs = self._rc_bufs = self._res_bufs = None

    def close(self):
        # (peer-mapped exchange buffers are NOT handed back here: close() may run from the garbage collector, at
        #  different moments on different ranks, and the pool must evolve identically everywhere)
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.pp_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # -- buffers -----------------------------------------------------------------------------
    def symbolic(self, st: structure.Structure, values_hint=None, cliques=None, comm=None):
        """``cliques``: (ptr, rows) of the nonzero border rows of the blocks of ALL ranks (several ranks only);
        ``comm``: the communicator that will reduce the exchange buffers (it may hand out peer-mapped ones)."""
        torch = self.torch
        self.st = st
        if cliques is not None:
            ptr = np.ascontiguousarray(cliques[0], dtype=np.int64)
            rows = np.ascontiguousarray(cliques[1], dtype=np.int32)
            self._check(self.lib.pp_set_coupling_cliques(self.handle, ptr.size - 1, native.np_ptr(ptr),
                                                         native.np_ptr(rows)), "pp_set_coupling_cliques")
        else:
            self._check(self.lib.pp_set_coupling_cliques(self.handle, 0, None, None), "pp_set_coupling_cliques")
        hint = None
        if values_hint is not None and values_hint.size == st.nvals and np.any(values_hint):
            hint = np.ascontiguousarray(values_hint, dtype=np.float64)
        code = self.lib.pp_symbolic(
            self.handle, st.n_local, native.np_ptr(st.block_n), native.np_ptr(st.border_ptr),
            native.np_ptr(st.border_rows), st.m_c, st.nvals, native.np_ptr(st.dest_front),
            native.np_ptr(st.dest_row), native.np_ptr(st.dest_col),
            native.np_ptr(hint) if hint is not None else None)
        if self._check(code, "pp_symbolic") != 0:
            return code
        mc = max(st.m_c, 1)
        self.schur_size = int(self.lib.pp_schur_size(self.handle))   # m_c^2, or the pattern size of a sparse S
        self._release_exchange()
        with torch.cuda.device(self.device):
            if comm is not None and comm.size > 1:
                self._comm = comm
                # the three exchange steps of the path; small ones may get peer-mapped double buffers
                self._schur_bufs = comm.exchange_buffers(max(self.schur_size, 1) + SCHUR_TAIL, self.device)
                self._rc_bufs = comm.exchange_buffers(mc, self.device)
                self._res_bufs = comm.exchange_buffers(mc + 2, self.device)
            else:
                self._schur_bufs = [torch.zeros(max(self.schur_size, 1) + SCHUR_TAIL, dtype=torch.float64, device=self.device)]
                self._rc_bufs = [torch.zeros(mc, dtype=torch.float64, device=self.device)]
                self._res_bufs = [torch.zeros(mc + 2, dtype=torch.float64, device=self.device)]
            self._turn = [0, 0, 0]
            self.schur, self.rc, self.resbuf = self._schur_bufs[0], self._rc_bufs[0], self._res_bufs[0]
        self.values_pin = torch.empty(max(st.nvals, 1), dtype=torch.float64, pin_memory=True)
        self.rhs_pin = torch.empty(max(st.local_dim, 1), dtype=torch.float64, pin_memory=True)
        self.x_pin = torch.empty(max(st.local_dim, 1), dtype=torch.float64, pin_memory=True)
        self.rhsc_pin = torch.empty(mc, dtype=torch.float64, pin_memory=True)
        self.xc_pin = torch.empty(mc, dtype=torch.float64, pin_memory=True)
        self.values = self.values_pin.numpy()
        return code

    # -- device-side regularisation ------------------------------------------------------------
    def set_classes(self, cls_local, cls_c):
        """Diagonal classes of the local rows / coupling rows (``pp_set_diagonal_classes``); before ``symbolic``."""
        if cls_local is None:
            self._check(self.lib.pp_set_diagonal_classes(self.handle, 0, None, 0, None), "pp_set_diagonal_classes")
            return
        a = np.ascontiguousarray(cls_local, dtype=np.int8)
        b = np.ascontiguousarray(cls_c, dtype=np.int8)
        self._check(self.lib.pp_set_diagonal_classes(self.handle, a.size, native.np_ptr(a), b.size, native.np_ptr(b)),
                    "pp_set_diagonal_classes")

    def set_shifts(self, shifts):
        arr = (C.c_double * 3)(*[float(v) for v in shifts])
        self._check(self.lib.pp_set_shifts(self.handle, arr), "pp_set_shifts")

    def value_uploads(self):
        return int(self.lib.pp_value_uploads(self.handle))

    def _next(self, which, bufs):
        """The exchange buffer to fill now: peer-mapped buffers come in pairs used alternately (a contribution may only
        be rewritten once every rank has passed the following exchange, see ``csrc/peer.cuh``)."""
        t = bufs[self._turn[which] % len(bufs)]
        self._turn[which] += 1
        return t

    # -- numeric -----------------------------------------------------------------------------
    def numeric_local(self, reuse=False):
        """Factor the local fronts from ``self.values`` (``reuse``: from the values of the previous call, which are
        still on the device -- a retry with new diagonal shifts); returns (status code, device S_local)."""
        self.schur = self._next(0, self._schur_bufs)
        if reuse:
            code = self.lib.pp_numeric_local(self.handle, None, 2, C.c_void_p(self.schur.data_ptr()), self._stream())
        else:
            code = self.lib.pp_numeric_local(self.handle, C.c_void_p(self.values_pin.data_ptr()), 0,
                                             C.c_void_p(self.schur.data_ptr()), self._stream())
        return self._check(code, "pp_numeric_local"), self.schur

    def schur_tail(self):
        out = (C.c_double * SCHUR_TAIL)()
        self._check(self.lib.pp_schur_tail(self.handle, out), "pp_schur_tail")
        return np.array(out[:], dtype=np.float64)

    def numeric_coupling(self, schur_sum):
        code = self.lib.pp_numeric_coupling(self.handle, C.c_void_p(schur_sum.data_ptr()), self._stream())
        return self._check(code, "pp_numeric_coupling")

    def inertia_local(self):
        out = (C.c_int64 * 3)()
        self._check(self.lib.pp_inertia_local(self.handle, out), "pp_inertia_local")
        return np.array(list(out), dtype=np.int64)

    def inertia_coupling(self):
        out = (C.c_int64 * 3)()
        self._check(self.lib.pp_inertia_coupling(self.handle, out), "pp_inertia_coupling")
        return np.array(list(out), dtype=np.int64)

    # -- solve -------------------------------------------------------------------------------
    def solve_forward(self):
        """Forward sweep on ``self.rhs_pin``; returns the device coupling contribution."""
        self.rc = self._next(1, self._rc_bufs)
        code = self.lib.pp_solve_forward(self.handle, C.c_void_p(self.rhs_pin.data_ptr()), 0,
                                         C.c_void_p(self.rc.data_ptr()), self._stream())
        self._check(code, "pp_solve_forward")
        return self.rc

    def solve_backward(self, rc_sum):
        code = self.lib.pp_solve_backward(self.handle, C.c_void_p(rc_sum.data_ptr()),
                                          C.c_void_p(self.rhsc_pin.data_ptr()), 0,
                                          C.c_void_p(self.x_pin.data_ptr()), C.c_void_p(self.xc_pin.data_ptr()),
                                          self._stream())
        self._check(code, "pp_solve_backward")
        return self.x_pin.numpy(), self.xc_pin.numpy()

    # -- iterative refinement (pp_residual_* / pp_refine_*) -------------------------------------
    def residual_local(self):
        """This rank's part of r = b - K x for the last solve; returns the device buffer
        ``[partial coupling rows (m_c) | |r_loc|^2 | |b_loc|^2]`` the caller sum-reduces."""
        self.resbuf = self._next(2, self._res_bufs)
        self._check(self.lib.pp_residual_local(self.handle, C.c_void_p(self.resbuf.data_ptr()), self._stream()),
                    "pp_residual_local")
        return self.resbuf

    def norms_ready(self):
        return self._auto_residual

    def residual_norms(self, buf_sum):
        out = (C.c_double * 2)()
        ptr = C.c_void_p(buf_sum.data_ptr()) if buf_sum is not None else None
        self._check(self.lib.pp_residual_norms(self.handle, ptr, out, self._stream()), "pp_residual_norms")
        return float(out[0]), float(out[1])

    def refine_forward(self):
        self.rc = self._next(1, self._rc_bufs)
        self._check(self.lib.pp_refine_forward(self.handle, C.c_void_p(self.rc.data_ptr()), self._stream()),
                    "pp_refine_forward")
        return self.rc

    def refine_backward(self, rc_sum, on_device=False):
        xp = None if on_device else C.c_void_p(self.x_pin.data_ptr())
        xcp = None if on_device else C.c_void_p(self.xc_pin.data_ptr())
        self._check(self.lib.pp_refine_backward(self.handle, C.c_void_p(rc_sum.data_ptr()), int(on_device), xp, xcp,
                                                self._stream()), "pp_refine_backward")
        return self.x_pin.numpy(), self.xc_pin.numpy()

    # -- device-resident variants (inputs / outputs already in HBM; used by bench.py `value`) ----
    def numeric_local_device(self, values_dev):
        self.schur = self._next(0, self._schur_bufs)
        code = self.lib.pp_numeric_local(self.handle, C.c_void_p(values_dev.data_ptr()), 1,
                                         C.c_void_p(self.schur.data_ptr()), self._stream())
        return self._check(code, "pp_numeric_local"), self.schur

    def solve_device(self, rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=None):
        self.rc = self._next(1, self._rc_bufs)
        self._check(self.lib.pp_solve_forward(self.handle, C.c_void_p(rhs_dev.data_ptr()), 1,
                                              C.c_void_p(self.rc.data_ptr()), self._stream()), "pp_solve_forward")
        rc = reduce(self.rc) if reduce is not None else self.rc
        self._check(self.lib.pp_solve_backward(self.handle, C.c_void_p(rc.data_ptr()),
                                               C.c_void_p(rhsc_dev.data_ptr()), 1, C.c_void_p(x_dev.data_ptr()),
                                               C.c_void_p(xc_dev.data_ptr()), self._stream()), "pp_solve_backward")

    def set_option(self, name, value):
        self._check(self.lib.pp_set_option(self.handle, name.encode(), float(value)), "pp_set_option")
        if name == "auto_residual":
            self._auto_residual = bool(value)

    def profile(self, reset=True):
        """Accumulated device milliseconds and launch counts per kernel class (see pp_profile)."""
        ms = (C.c_double * 8)()
        cnt = (C.c_int64 * 8)()
        self._check(self.lib.pp_profile(self.handle, ms, cnt, 1 if reset else 0), "pp_profile")
        names = ("assemble", "panel", "swaps", "update", "schur", "forward", "backward", "subtree")
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(names)}

    def coupling_stats(self):
        out = (C.c_int64 * 8)()
        self._check(self.lib.pp_coupling_stats(self.handle, out), "pp_coupling_stats")
        keys = ("levels", "schur_size", "m_c", "bottom_m_c", "blocks", "max_front", "level1_blocks")
        return dict(zip(keys, [int(v) for v in out]))

    def plan_stats(self, block=0):
        out = (C.c_int64 * 12)()
        self._check(self.lib.pp_plan_stats(self.handle, block, out), "pp_plan_stats")
        keys = ("plan", "supernodes", "root_cols", "delay_slots", "nnz_l_subtree", "max_front", "factor_doubles",
                "stack_doubles", "n_plans", "fell_back_dense", "delayed_to_root", "failed")
        return dict(zip(keys, [int(v) for v in out]))

    def int_tensor(self, values):
        t = self.ints[: len(values)]
        t.copy_(self.torch.tensor(list(values), dtype=self.torch.int64))
        return t

    def kernel_launches(self):
        return int(self.lib.pp_kernel_launches(self.handle))

    def factor_bytes(self):
        return int(self.lib.pp_factor_bytes(self.handle))


class B200SchurComplementLinearSolver(LinearSolverInterface):
    """Solve ``K x = b`` for block-bordered-diagonal symmetric ``K``::

          K1          A1^T
              K2      A2^T
                  K3  A3^T
          A1  A2  A3  Q

    on one B200 per process.  Only the lower border ``A_i`` and the lower triangles of ``K_i`` and
    ``Q`` are read (as the MA27 / MUMPS leaves do).

    Parameters
    ----------
    subproblem_solvers, schur_complement_solver:
        accepted for signature compatibility with the classes being replaced
        (``explicit_schur_complement.py:28-29``) and ignored -- this solver owns its leaves.
    device: CUDA device index (default: current device).
    comm: ``Communicator`` (default: the ``torch.distributed`` world if initialised, else 1 rank).
    options: native tunables, e.g. ``{"pivot_tol": 0.0, "panel_width": 64}``.
    regularization_classes: ``interface.regularization_classes()`` of a ``DeviceRegularizationMixin`` interface, so
        that the diagonal entries the inertia-correction loop shifts are reserved by the first symbolic phase.
    """

    def __init__(self, subproblem_solvers=None, schur_complement_solver=None, device=None, comm=None,
                 options=None, backend=None, refine_tol=2e-11, max_refine=2, regularization_classes=None):
        self.subproblem_solvers = subproblem_solvers
        # (per diagonal block, coupling) diagonal classes for device-side inertia correction (regularization.py);
        # learnt from the first RegularizedKKT otherwise (at the price of one more symbolic phase)
        self._classes = regularization_classes
        self._uploaded_token = None
        self.refine_tol = float(refine_tol)
        self.max_refine = int(max_refine)
        self.last_residual = None
        self.refine_steps = 0
        self._copiers = [None, None, None]
        env = os.environ.get("PARAPINT_B200_HOST_THREADS")   # 0 disables the threaded host gather
        self.host_threads = int(env) if env not in (None, "") else None
        self.schur_complement_solver = schur_complement_solver
        self.comm = comm if comm is not None else Communicator()
        self.backend = backend if backend is not None else CudaBackend(device, options)
        # one rank: no collective between the local and the coupling phase, status + inertia are read with one host
        # sync (defer_status 1; 0 on request).  Several ranks: always 2 -- the local phase is not synchronised at all,
        # its status and overflow flag travel in the tail of the Schur all-reduce (a per-rank status read would let
        # ranks disagree and desynchronise the collectives).
        want = (options or {}).get("defer_status")
        if self.comm.size > 1:
            if want not in (None, 2):
                raise ValueError("defer_status must be 2 (or unset) when the communicator has more than one rank")
            self._defer = 2
        else:
            self._defer = 1 if want is None else int(want)
            if self._defer not in (0, 1):
                raise ValueError("defer_status must be 0 or 1 on a single rank")
        self.backend.set_option("defer_status", self._defer)
        if self.comm.size == 1 and self._defer == 1 and refine_tol > 0 and max_refine > 0 \
                and isinstance(self.backend, CudaBackend):
            # nothing to reduce: the residual norms of the first solve ride on the copy-out's synchronisation
            self.backend.set_option("auto_residual", 1)
        self.symbolic_calls = 0
        self.block_dim = 0
        self.block_matrix = None
        self.local_block_indices = []
        self._st = None
        self._status = None
        self._tail = None
        self.logger = self.getLogger()

    @classmethod
    def getLoggerName(cls):
        return "b200_schur"

    # ---------------------------------------------------------------------------------------
    def _result(self, code, raise_on_error, what):
        res = LinearSolverResults()
        res.status = LinearSolverStatus(int(code))
        self._status = res.status
        if res.status not in _OK and raise_on_error:
            detail = getattr(self.backend, "last_error", "")
            raise RuntimeError(f"{what} unsuccessful; status: {res.status}" + (f" ({detail})" if detail else ""))
        return res

    def _finish(self, code, raise_on_error, what):
        """Agree on the worst status across ranks (``_gather_results``, ``mpi...:19-30``): the codes are ranked
        successful < warning < not_enough_memory < singular < error before the MAX reduction (the raw enum values do
        not order that way), so one rank's error is never masked by another rank's warning."""
        if self.comm.size > 1:
            t = self.backend.int_tensor([_SEVERITY[int(code)]])
            self.comm.allreduce_max_(t)
            code = _FROM_SEVERITY[int(t[0].item())]
        return self._result(code, raise_on_error, what)

    def _analyse(self, matrix):
        """Structure + native symbolic phase; returns this rank's status code (agreed by the caller)."""
        st = structure.analyse(matrix, self.comm.rank, self.comm.size)
        self.block_dim = st.n_blocks + 1
        self.local_block_indices = list(st.local_blocks)
        hint = np.zeros(st.nvals)
        structure.gather_values(matrix, st, hint)  # values only steer the ordering (2x2 pivot pre-selection)
        self._copiers = [None, None, None]               # pointer tables of the old structure are stale
        self._uploaded_token = None
        if self._classes is not None:
            per_block, coupling = self._classes
            parts = [np.asarray(per_block[i], dtype=np.int8) for i in st.local_blocks]
            for f, (i, part) in enumerate(zip(st.local_blocks, parts)):
                if part.size != st.block_n[f]:
                    raise ValueError(f"regularization classes of block {i} have size {part.size}, expected {st.block_n[f]}")
            cls_c = np.asarray(coupling, dtype=np.int8)
            if cls_c.size != st.m_c:
                raise ValueError(f"regularization classes of the coupling rows have size {cls_c.size}, expected {st.m_c}")
            self.backend.set_classes(np.concatenate(parts) if parts else np.zeros(0, dtype=np.int8), cls_c)
        cliques = None
        if self.comm.size > 1:
            # the pattern of S is the union over the blocks of EVERY rank (mpi...:244-247 all-gathers the same lists)
            parts = self.comm.allgather_object((np.diff(st.border_ptr), st.border_rows))
            lens = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, dtype=np.int64)
            cliques = (np.concatenate(([0], np.cumsum(lens))), np.concatenate([p[1] for p in parts]))
        code = self.backend.symbolic(st, hint, cliques, self.comm)
        self._st = st
        self._status = None
        self.symbolic_calls += 1
        return code

    def do_symbolic_factorization(self, matrix, raise_on_error=True, timer=None):
        """Analyse the block structure and allocate the fronts (``explicit...:44-78``,
        ``mpi...:165-255``).  Collective over the communicator."""
        timer = timer or _NullTimer()
        timer.start("sc_structure")
        if isinstance(matrix, RegularizedKKT):
            if self._classes is None:
                self._classes = matrix.classes
            matrix = matrix.base
        code = self._analyse(matrix)
        timer.stop("sc_structure")
        return self._finish(code, raise_on_error, "Symbolic factorization")

    def do_numeric_factorization(self, matrix, raise_on_error=True, timer=None):
        """Factor every local front, all-reduce the Schur complement, factor it
        (``explicit...:80-129``, ``mpi...:257-361``).  Collective."""
        if self._st is None:
            raise RuntimeError("do_symbolic_factorization must be called before do_numeric_factorization")
        timer = timer or _NullTimer()
        shifts, token = (0.0, 0.0, 0.0), getattr(matrix, "_pp_token", None)
        if isinstance(matrix, RegularizedKKT):
            # base matrix + diagonal shifts (regularization.py): the shifts are applied on the device, and when the
            # values of `base` are the ones already there (same evaluation) nothing is gathered or uploaded
            shifts, token = matrix.shifts, matrix.token
            if self._classes is None:
                # first regularisation and the classes were not given at construction: reserve the diagonal entries
                # now (one more symbolic phase; every rank is here together -- the IPM loop is SPMD)
                self._classes = matrix.classes
                sym = self._finish(self._analyse(matrix.base), False, "Symbolic factorization")
                if sym.status not in _OK:
                    return self._result(sym.status.value, raise_on_error, "Numeric factorization")
            matrix = matrix.base
        self.block_matrix = matrix
        be = self.backend
        timer.start("form SC")
        reuse = token is not None and token == self._uploaded_token
        if not reuse:
            self._wake(0)   # the copy pool wakes up while the matrix is being walked
        changed = False if reuse else not structure.gather_values(matrix, self._st, be.values, self._copier(0))
        self._uploaded_token = None
        has_shifts = self._classes is not None or any(shifts)
        if self.comm.size == 1:
            if changed:
                # COO pattern / ordering changed since the symbolic phase (happens after the first
                # regularisation): analyse again, as mumps_interface.py:82-83 does.
                self.logger.debug("nonzero pattern changed; repeating the symbolic phase")
                code = self._analyse(matrix)
                if code != 0:
                    timer.stop("form SC")
                    return self._result(code, raise_on_error, "Numeric factorization")
                if not structure.gather_values(matrix, self._st, be.values, self._copier(0)):
                    raise RuntimeError("could not gather the matrix values after re-analysis")
            timer.start("factorize")
            if has_shifts:
                be.set_shifts(shifts)
            code, schur_local = be.numeric_local(reuse and not changed)
            timer.stop("factorize")
            self._tail = None
            if code in (0, LinearSolverStatus.singular.value):
                self._uploaded_token = token   # the values of this evaluation are on the device
            if code != 0:   # not deferred (or a run-time failure): nothing to factor in the coupling phase
                timer.stop("form SC")
                return self._result(code, raise_on_error, "Numeric factorization")
            timer.stop("form SC")
            timer.start("factor SC")
            self._wake(1)   # ... and stays awake through the wait for the GPU: the right-hand side is packed next
            code = be.numeric_coupling(schur_local)
            timer.stop("factor SC")
            return self._result(code, raise_on_error, "Numeric factorization")

        # Several ranks.  ONE collective and ONE host synchronisation per factorisation in the common case: the local
        # phase is not synchronised, the tail of the reduced Schur buffer carries every rank's status, the summed
        # inertia (mpi...:21,343,427-429) and two "do it again" requests, and the coupling phase reads it:
        #   tail[0] some block singular      tail[1] some rank's sparse path overflowed (redo synchronously, dense)
        #   tail[2:5] inertia                tail[5] some rank's nonzero pattern changed (re-analyse, collectively)
        #   NaN anywhere = some rank failed at run time (CUDA error, capacity): every rank returns `error`.
        sync, reanalysed = False, False
        code = LinearSolverStatus.error.value
        for _ in range(4):
            n_s = be.schur_size
            if changed:
                schur_local = be.schur = be._next(0, be._schur_bufs) if hasattr(be, "_next") else be.schur
                schur_local.zero_()
                schur_local[n_s + 5] = 1.0
                local_code = -1                       # this rank has nothing factorised
            else:
                timer.start("factorize")
                if sync:
                    be.set_option("defer_status", 0)
                if has_shifts:
                    be.set_shifts(shifts)
                try:
                    local_code, schur_local = be.numeric_local(reuse)
                finally:
                    if sync:
                        be.set_option("defer_status", 2)
                timer.stop("factorize")
                if local_code in (LinearSolverStatus.error.value, LinearSolverStatus.not_enough_memory.value):
                    schur_local[n_s:n_s + SCHUR_TAIL] = float("nan")   # seen by every rank after the reduction
            timer.start("communicate")
            schur_local = self.comm.allreduce_sum_(schur_local)
            timer.stop("communicate")
            if local_code in (0, LinearSolverStatus.singular.value):
                self._wake(1)
                code = be.numeric_coupling(schur_local)   # reads the reduced tail: same answer on every rank
                tail = be.schur_tail()
            else:
                tail = schur_local[n_s:n_s + SCHUR_TAIL].cpu().numpy()
                code = LinearSolverStatus.error.value
            self._tail = tail
            if not np.all(np.isfinite(tail)):
                code = LinearSolverStatus.error.value
                break
            if tail[5] > 0:
                if reanalysed:
                    code = LinearSolverStatus.error.value
                    break
                self.logger.debug("nonzero pattern changed on some rank; repeating the symbolic phase on all")
                reanalysed = True
                sym = self._finish(self._analyse(matrix), False, "Symbolic factorization")
                if sym.status not in _OK:
                    code = sym.status.value
                    break
                be = self.backend
                reuse = False
                changed = not structure.gather_values(matrix, self._st, be.values, self._copier(0))
                continue
            if tail[1] > 0:
                if sync:
                    code = LinearSolverStatus.error.value
                    break
                sync = True                            # the overflowing rank re-analyses densely
                continue
            break
        if code in (0, LinearSolverStatus.singular.value):
            self._uploaded_token = token
        timer.stop("form SC")
        return self._result(code, raise_on_error, "Numeric factorization")

    def do_back_solve(self, rhs, timer=None):
        """Three-phase block elimination (``explicit...:131-155``, ``mpi...:363-402``); returns a new
        vector with the block structure of ``rhs``; ``rhs`` is not modified.  Collective."""
        if self._status not in _OK or self.block_matrix is None:
            raise RuntimeError("do_numeric_factorization must succeed before do_back_solve")
        timer = timer or _NullTimer()
        timer.start("back_solve")
        st = self._st
        structure.pack_rhs(rhs, st, self.backend.rhs_pin.numpy(), self._copier(1))
        self.backend.rhsc_pin.numpy()[: st.m_c] = structure.coupling_rhs(rhs, st)
        be = self.backend
        be.failed = False
        rc = be.solve_forward()
        if be.failed:
            rc.fill_(float("nan"))      # a run-time failure on this rank: every rank sees it after the reduction
        rc = self.comm.allreduce_sum_(rc)
        self._wake(2)   # the solution is scattered by the copy pool right after the wait for the GPU
        x_local, x_c = be.solve_backward(rc)
        if be.failed and self.comm.size == 1:
            raise RuntimeError(f"back solve failed: {be.last_error}")
        x_local, x_c = self._refine(x_local, x_c)
        if be.failed or not np.all(np.isfinite(x_c[: st.m_c])):
            raise RuntimeError("back solve failed" + (f": {be.last_error}" if be.last_error else ""))
        out = structure.unpack_solution(rhs, st, x_local, x_c[: st.m_c], self._copier(2))
        timer.stop("back_solve")
        return out

    def _copier(self, which):
        """Threaded host gather / scatter for the CUDA backend (values: slot 0, right-hand side: slot 1, solution:
        slot 2; one pointer-table cache each)."""
        if not isinstance(self.backend, CudaBackend) or self.host_threads == 0:
            return None
        if self._copiers[which] is None:
            self._copiers[which] = native.HostCopier(self.host_threads)
            if which == 0:
                self._copiers[0].stage = (self.backend.handle, self.backend._stream)
        return self._copiers[which]

    def _wake(self, which):
        """Announce a host copy to the pool behind the copiers (``pp_host_wake``; bounded spin, see native.HostCopier)."""
        copier = self._copier(which)
        if copier is not None:
            copier.wake()

    def _refine(self, x_local, x_c):
        """Iterative refinement with the values that were factorised: while ||b - K x|| > refine_tol ||b||,
        solve K d = r with the same factors and add d.  One extra sum-reduction of m_c + 2 doubles per
        residual and one of m_c per correction; every rank takes the same branch (the norms are global)."""
        self.refine_steps = 0
        self.last_residual = None
        if self.max_refine <= 0 or self.refine_tol <= 0:
            return x_local, x_c
        be = self.backend
        for step in range(self.max_refine + 1):
            if step == 0 and getattr(be, "norms_ready", lambda: False)():
                r2, b2 = be.residual_norms(None)            # formed by pp_solve_backward ("auto_residual")
            elif getattr(be, "failed", False):
                # this rank's solve failed at run time: keep the collectives aligned, every rank sees NaN norms
                buf = be.resbuf = be._next(2, be._res_bufs) if hasattr(be, "_next") else be.resbuf
                buf.fill_(float("nan"))
                self.comm.allreduce_sum_(buf)
                r2 = b2 = float("nan")
            else:
                buf = self.comm.allreduce_sum_(be.residual_local())
                r2, b2 = be.residual_norms(buf)
            rel = float(np.sqrt(r2 / b2)) if b2 > 0 else float(np.sqrt(r2))
            stalled = self.last_residual is not None and not rel < 0.25 * self.last_residual
            self.last_residual = rel if self.last_residual is None else min(rel, self.last_residual)
            if stalled or not np.isfinite(rel) or rel <= self.refine_tol or step == self.max_refine:
                break                                   # converged, or at the floor eps*|K||x|/|b| of this system
            rc = self.comm.allreduce_sum_(be.refine_forward())
            x_local, x_c = be.refine_backward(rc)
            self.refine_steps += 1
        return x_local, x_c

    def get_inertia(self):
        """(num_pos, num_neg, num_zero) summed over all blocks and the Schur complement
        (Haynsworth additivity; ``explicit...:157-172``, ``mpi...:404-436``).  Collective."""
        if self._status not in _OK or self.block_matrix is None:
            raise RuntimeError("The inertia is only available after a successful do_numeric_factorization.")
        if self.comm.size > 1 and self._tail is not None:
            local = np.rint(self._tail[2:5]).astype(np.int64)  # summed over ranks by the Schur all-reduce
        else:
            local = self.backend.inertia_local()
        tot = local + self.backend.inertia_coupling()
        return int(tot[0]), int(tot[1]), int(tot[2])

    def increase_memory_allocation(self, factor):
        """Factor storage is sized exactly in the symbolic phase (a block that outgrows its delayed-pivot capacity
        is re-analysed as dense fronts inside the numeric phase), so there is no work array to grow
        (``explicit...:174-177`` forwards to the leaves).  ``not_enough_memory`` is only reported when an allocation
        itself fails (host ``bad_alloc`` / ``cudaMalloc``), which a growth factor cannot cure."""
        return None

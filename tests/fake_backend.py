"""CPU stand-in for ``parapint_b200.schur_solver.CudaBackend`` (TEST INFRASTRUCTURE).

Implements the contract of the C ABI (include/parapint_b200.h) with numpy so that the *host-side*
logic of ``B200SchurComplementLinearSolver`` -- partitioning, value gathering, the order of the
collectives, solution unpacking, status / inertia agreement -- can be exercised on CPU, including with
``torch.distributed`` (gloo) at world_size > 1.  It is never importable from the product package.
"""
import numpy as np
import torch

from oracle.schur_oracle import dense_inertia


class FakeBackend:
    def __init__(self):
        self.launches = 0
        self.last_error = ""
        self.failed = False
        self.fail_symbolic = self.fail_numeric = False   # test hooks: a run-time failure on this rank
        self.ints = torch.zeros(4, dtype=torch.int64)
        self.cls_local = self.cls_c = None
        self.shifts = np.zeros(4)
        self.uploads = 0
        self.device_values = None

    def set_option(self, name, value):
        pass

    # device-side regularisation (pp_set_diagonal_classes / pp_set_shifts / PP_VALUES_REUSE)
    def set_classes(self, cls_local, cls_c):
        self.cls_local = None if cls_local is None else np.asarray(cls_local, dtype=np.int64)
        self.cls_c = None if cls_c is None else np.asarray(cls_c, dtype=np.int64)
        self.shifts = np.zeros(4)

    def set_shifts(self, shifts):
        self.shifts = np.concatenate(([0.0], np.asarray(shifts, dtype=float)))

    def value_uploads(self):
        return self.uploads

    def symbolic(self, st, values_hint=None, cliques=None, comm=None):
        self.st = st
        self.cliques = cliques
        if self.fail_symbolic:
            self.last_error = "injected symbolic failure"
            return 3
        mc = max(st.m_c, 1)
        self.schur_size = st.m_c * st.m_c
        self.schur = torch.zeros(max(self.schur_size, 1) + 8, dtype=torch.float64)
        self.rc = torch.zeros(mc, dtype=torch.float64)
        self.ints = torch.zeros(4, dtype=torch.int64)
        self.device_values = None
        self.values_pin = torch.zeros(max(st.nvals, 1), dtype=torch.float64)
        self.rhs_pin = torch.zeros(max(st.local_dim, 1), dtype=torch.float64)
        self.x_pin = torch.zeros(max(st.local_dim, 1), dtype=torch.float64)
        self.rhsc_pin = torch.zeros(mc, dtype=torch.float64)
        self.xc_pin = torch.zeros(mc, dtype=torch.float64)
        self.values = self.values_pin.numpy()
        return 0

    def _fronts(self):
        st = self.st
        sizes = [int(st.block_n[f] + st.border_ptr[f + 1] - st.border_ptr[f]) for f in range(st.n_local)] + [st.m_c]
        fronts = [np.zeros((s, s)) for s in sizes]
        for k in range(st.nvals):
            f = st.dest_front[k]
            if f >= 0:
                fronts[f][st.dest_row[k], st.dest_col[k]] += self.device_values[k]
        if self.cls_local is not None:
            for f in range(st.n_local):
                n = int(st.block_n[f])
                d = self.shifts[self.cls_local[st.rhs_offsets[f]:st.rhs_offsets[f + 1]]]
                fronts[f][np.arange(n), np.arange(n)] += d
            fronts[-1][np.arange(st.m_c), np.arange(st.m_c)] += self.shifts[self.cls_c]
        return [np.tril(F) + np.tril(F, -1).T for F in fronts]

    def numeric_local(self, reuse=False):
        st = self.st
        if not reuse:
            self.device_values = self.values.copy()
            self.uploads += 1
        assert self.device_values is not None
        if self.fail_numeric:
            self.last_error = "injected numeric failure"
            return 3, self.schur
        fronts = self._fronts()
        self.K, self.A, self.rows, self.inert = [], [], [], np.zeros(3, dtype=np.int64)
        S = np.zeros((st.m_c, st.m_c))
        code = 0
        for f in range(st.n_local):
            n = int(st.block_n[f])
            K, A = fronts[f][:n, :n], fronts[f][n:, :n]
            rows = st.border_rows[st.border_ptr[f]:st.border_ptr[f + 1]]
            self.K.append(K); self.A.append(A); self.rows.append(rows)
            if n and np.linalg.matrix_rank(K) < n:
                code = 2
                continue
            self.inert += np.asarray(dense_inertia(K, "eigvalsh"), dtype=np.int64)
            if rows.size:
                S[np.ix_(rows, rows)] -= A @ np.linalg.solve(K, A.T)
        self.Q = fronts[-1]
        tail = np.array([1.0 if code == 2 else 0.0, 0.0, *self.inert.astype(float), 0.0, 0.0, 0.0])
        self.schur.zero_()
        self.schur[: self.schur_size] = torch.from_numpy(S.T.reshape(-1).copy())
        self.schur[self.schur_size: self.schur_size + 8] = torch.from_numpy(tail)
        return code, self.schur

    def numeric_coupling(self, schur_sum):
        """Like pp_numeric_coupling with defer_status = 2: reads the (reduced) tail of the Schur buffer."""
        mc = self.st.m_c
        self.tail = schur_sum.numpy()[self.schur_size: self.schur_size + 8].copy()
        if not np.all(np.isfinite(self.tail)):
            return 3
        if self.tail[1] > 0:
            return 1
        if self.tail[0] > 0:
            return 2
        self.S = self.Q + schur_sum.numpy()[: mc * mc].reshape(mc, mc).T
        if mc and np.linalg.matrix_rank(self.S) < mc:
            return 2
        self.inert_c = np.asarray(dense_inertia(self.S, "eigvalsh"), dtype=np.int64)
        return 0

    def schur_tail(self):
        return self.tail.copy()

    def inertia_local(self):
        return self.inert.copy()

    def inertia_coupling(self):
        return self.inert_c.copy()

    def solve_forward(self):
        st = self.st
        rc = np.zeros(st.m_c)
        self.y = []
        for f in range(st.n_local):
            r = self.rhs_pin.numpy()[st.rhs_offsets[f]:st.rhs_offsets[f + 1]]
            y = np.linalg.solve(self.K[f], r)
            self.y.append(y)
            if self.rows[f].size:
                rc[self.rows[f]] -= self.A[f] @ y
        self.rc[: st.m_c] = torch.from_numpy(rc)
        return self.rc

    def solve_backward(self, rc_sum):
        st = self.st
        xc = np.linalg.solve(self.S, self.rhsc_pin.numpy()[: st.m_c] + rc_sum.numpy()[: st.m_c]) if st.m_c else np.zeros(0)
        self.xc_pin.numpy()[: st.m_c] = xc
        for f in range(st.n_local):
            r = self.rhs_pin.numpy()[st.rhs_offsets[f]:st.rhs_offsets[f + 1]]
            corr = self.A[f].T @ xc[self.rows[f]] if self.rows[f].size else 0.0
            self.x_pin.numpy()[st.rhs_offsets[f]:st.rhs_offsets[f + 1]] = np.linalg.solve(self.K[f], r - corr)
        return self.x_pin.numpy(), self.xc_pin.numpy()

    # -- iterative refinement --------------------------------------------------------------
    def _matvec_parts(self):
        st = self.st
        x, xc = self.x_pin.numpy(), self.xc_pin.numpy()[: st.m_c]
        b = self.rhs_pin.numpy()
        r = np.zeros(max(st.local_dim, 1))
        part = np.zeros(st.m_c)
        for f in range(st.n_local):
            sl = slice(st.rhs_offsets[f], st.rhs_offsets[f + 1])
            r[sl] = b[sl] - self.K[f] @ x[sl]
            if self.rows[f].size:
                r[sl] -= self.A[f].T @ xc[self.rows[f]]
                part[self.rows[f]] -= self.A[f] @ x[sl]
        return r, part

    def residual_local(self):
        st = self.st
        r, part = self._matvec_parts()
        self.r_loc = r
        b = self.rhs_pin.numpy()[: st.local_dim]
        self.resbuf = torch.from_numpy(np.concatenate([part, [float(r[: st.local_dim] @ r[: st.local_dim]), float(b @ b)]]))
        return self.resbuf

    def residual_norms(self, buf_sum):
        st = self.st
        buf = buf_sum.numpy()
        bc, xc = self.rhsc_pin.numpy()[: st.m_c], self.xc_pin.numpy()[: st.m_c]
        self.r_c = bc - self.Q @ xc + buf[: st.m_c]
        return float(buf[st.m_c] + self.r_c @ self.r_c), float(buf[st.m_c + 1] + bc @ bc)

    def refine_forward(self):
        st = self.st
        rc = np.zeros(st.m_c)
        for f in range(st.n_local):
            y = np.linalg.solve(self.K[f], self.r_loc[st.rhs_offsets[f]:st.rhs_offsets[f + 1]])
            if self.rows[f].size:
                rc[self.rows[f]] -= self.A[f] @ y
        self.rc[: st.m_c] = torch.from_numpy(rc)
        return self.rc

    def refine_backward(self, rc_sum, on_device=False):
        st = self.st
        dc = np.linalg.solve(self.S, self.r_c + rc_sum.numpy()[: st.m_c]) if st.m_c else np.zeros(0)
        self.xc_pin.numpy()[: st.m_c] += dc
        for f in range(st.n_local):
            sl = slice(st.rhs_offsets[f], st.rhs_offsets[f + 1])
            corr = self.A[f].T @ dc[self.rows[f]] if self.rows[f].size else 0.0
            self.x_pin.numpy()[sl] += np.linalg.solve(self.K[f], self.r_loc[sl] - corr)
        return self.x_pin.numpy(), self.xc_pin.numpy()

    def int_tensor(self, values):
        t = self.ints[: len(values)]
        t.copy_(torch.tensor(list(values), dtype=torch.int64))
        return t

    def kernel_launches(self):
        return 0

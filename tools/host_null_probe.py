"""Host side of an e2e step with the GPU calls stubbed out (runs without a GPU): what the interpreter, the value gather,
the right-hand-side packing and the scattering of the solution cost per step at BASELINE config 2.
`python tools/host_null_probe.py [fresh]` -- `fresh`: newly built matrix / vector objects every step."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.kkt_generator import EstimationModel  # noqa: E402
from parapint_b200 import schur_solver  # noqa: E402
from parapint_b200.schur_solver import B200SchurComplementLinearSolver, CudaBackend  # noqa: E402


class _Buf:
    """numpy array with the two tensor methods the solver touches."""

    def __init__(self, n):
        self.a = np.zeros(max(n, 1))

    def numpy(self):
        return self.a

    def data_ptr(self):
        return self.a.ctypes.data


class NullBackend(CudaBackend):
    def __init__(self):
        self.last_error, self.failed, self._auto_residual = "", False, True
        self.handle = None

    def symbolic(self, st, values_hint=None, cliques=None, comm=None):
        self.st = st
        self.schur_size = st.m_c * st.m_c
        self.values_pin, self.rhs_pin, self.x_pin = _Buf(st.nvals), _Buf(st.local_dim), _Buf(st.local_dim)
        self.rhsc_pin, self.xc_pin = _Buf(st.m_c), _Buf(st.m_c)
        self.values = self.values_pin.numpy()
        self.schur = self.rc = self.resbuf = None
        return 0

    def close(self): pass
    def set_option(self, name, value): pass
    def set_classes(self, a, b): pass
    def set_shifts(self, s): pass
    def numeric_local(self, reuse=False): return 0, None
    def numeric_coupling(self, s): return 0
    def inertia_local(self): return np.array([1, 1, 0])
    def inertia_coupling(self): return np.array([1, 1, 0])
    def solve_forward(self): return None
    def solve_backward(self, rc): return self.x_pin.numpy(), self.xc_pin.numpy()
    def norms_ready(self): return True
    def residual_norms(self, buf): return 0.0, 1.0


def main():
    fresh = len(sys.argv) > 1 and sys.argv[1] == "fresh"
    m = EstimationModel(64, 150, 6, 50)
    pool = [(m.build_kkt(), m.build_rhs()) for _ in range(6 if fresh else 1)]
    s = B200SchurComplementLinearSolver(backend=NullBackend())
    s.backend.__class__ = NullBackend
    s.do_symbolic_factorization(pool[0][0])
    for c in range(3):                       # the copiers exist from the first step on; no staging without a device
        cp = s._copier(c)
        if cp is not None:
            cp.stage = None

    def step(k):
        kkt, rhs = pool[k % len(pool)]
        s.do_numeric_factorization(kkt)
        s.get_inertia()
        return s.do_back_solve(rhs)

    for k in range(20):
        step(k)
    reps = 300
    t0 = time.perf_counter()
    for k in range(reps):
        step(k)
    print(f"host side per step: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms ({'fresh' if fresh else 'same'} objects, "
          f"{s._copier(0).threads if s._copier(0) else 0} copy threads)")
    pr = cProfile.Profile()
    pr.enable()
    for k in range(reps):
        step(k)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()

"""Wall-clock phases of one e2e step (config 2) through the plugin: where the host time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver, structure
import parapint_b200.schur_solver as SS
m = EstimationModel(64, 150, 6, 50); kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver(); s.do_symbolic_factorization(kkt)
acc = {}
def wrap(obj, name, label=None):
    fn = getattr(obj, name); label = label or name
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, w)
for n in ("gather_values", "pack_rhs", "unpack_solution", "coupling_rhs"): wrap(structure, n)
be = s.backend
for n in ("numeric_local", "numeric_coupling", "solve_forward", "solve_backward", "residual_norms", "inertia_local", "inertia_coupling", "set_shifts"): wrap(be, n)
wrap(s, "_refine")
def step():
    s.do_numeric_factorization(kkt); s.get_inertia(); return s.do_back_solve(rhs)
for _ in range(20): step()
acc.clear(); torch.cuda.synchronize(); N = 200
t0 = time.perf_counter()
for _ in range(N): step()
tot = (time.perf_counter() - t0) / N * 1e3
print("ms/step", round(tot, 4))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]): print(f"  {k:20s} {v / N * 1e3:.4f} ms")
print("  (unaccounted)       ", round(tot - sum(v for k, v in acc.items() if k != "residual_norms") / N * 1e3, 4))

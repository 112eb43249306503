"""Quick timing + plan statistics on a generator workload (not the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver

args = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else [64, 150, 6, 50]
opts = {"profile": 1}
for a in sys.argv[5:]:
    k, v = a.split("=")
    opts[k] = float(v)
m = EstimationModel(*args)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver(options=opts)
t0 = time.perf_counter(); s.do_symbolic_factorization(kkt); torch.cuda.synchronize(); print("symbolic s", time.perf_counter() - t0)
for rep in range(3):
    t0 = time.perf_counter()
    st = s.do_numeric_factorization(kkt).status
    ine = s.get_inertia()
    x = s.do_back_solve(rhs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("rep", rep, "ms", dt * 1e3, st, ine == m.expected_inertia(), "max_err", m.check_result(x))
print(s.backend.plan_stats(0))
print({k: (round(v["ms"] / 3, 3), v["launches"] // 3) for k, v in s.backend.profile().items()})
from oracle.schur_oracle import sym_full
K = sym_full(kkt); b = rhs.flatten()
print("rel residual", np.linalg.norm(K @ x.flatten() - b) / np.linalg.norm(b))

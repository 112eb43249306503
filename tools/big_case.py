"""Large generator cases (config-5-like wide coupling): timing + correctness vs closed-form inertia and residual."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
args = [int(a) for a in sys.argv[1:5]]
check_oracle = len(sys.argv) > 5 and sys.argv[5] == "oracle"
t0 = time.perf_counter(); m = EstimationModel(*args); kkt, rhs = m.build_kkt(), m.build_rhs(); print("build model s", time.perf_counter() - t0, "block n", m.block_dim)
opts = {"profile": 1}
for a in sys.argv[5:]:
    if "=" in a:
        k, v = a.split("="); opts[k] = float(v)
s = B200SchurComplementLinearSolver(options=opts)
t0 = time.perf_counter(); s.do_symbolic_factorization(kkt); torch.cuda.synchronize(); print("symbolic s", time.perf_counter() - t0)
print(s.backend.plan_stats(0), "factor GB", s.backend.factor_bytes() / 1e9)
for rep in range(3):
    t0 = time.perf_counter(); st = s.do_numeric_factorization(kkt).status; torch.cuda.synchronize(); t1 = time.perf_counter()
    ine = s.get_inertia(); x = s.do_back_solve(rhs); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep", rep, "numeric ms", (t1 - t0) * 1e3, "solve ms", (t2 - t1) * 1e3, st, ine == m.expected_inertia(), "max_err", m.check_result(x))
print({k: (round(v["ms"] / 3, 3), v["launches"] // 3) for k, v in s.backend.profile().items()})
N = args[0]; border = m.border().tocsr(); xc = np.asarray(x.get_block(N)); num = den = 0.0; cres = np.zeros(args[3])
for i in range(N):
    K = kkt.get_block(i, i).tocsr(); xi = np.asarray(x.get_block(i)); ri = np.asarray(rhs.get_block(i))
    num += np.sum((K @ xi + border.T @ xc - ri) ** 2); den += np.sum(ri ** 2); cres += border @ xi
print("rel residual", np.sqrt(num + np.sum(cres ** 2)) / np.sqrt(den))
if check_oracle:
    from oracle.schur_oracle import SchurOracle
    o = SchurOracle(); o.symbolic(kkt); t0 = time.perf_counter(); o.numeric(kkt); xr = o.solve(rhs); print("oracle s", time.perf_counter() - t0)
    print("rel diff vs oracle", np.linalg.norm(x.flatten() - xr.flatten()) / np.linalg.norm(xr.flatten()))

"""Config-4-shaped blocks (family P): n_x=10000, n_eq=8000, n_in=1000, n_fs=200 -> 20200 rows per block."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import stochastic_ipm_system, block_vector
from parapint_b200 import B200SchurComplementLinearSolver
from oracle.schur_oracle import sym_full
nb = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
opts = {"profile": 1}
same = False
for a in sys.argv[3:]:
    k, v = a.split("=")
    if k == "same": same = bool(int(v))
    else: opts[k] = float(v)
n_x, n_eq, n_in, n_fs = int(10000 * scale), int(8000 * scale), int(1000 * scale), int(200 * min(1.0, scale * 2))
t0 = time.perf_counter(); kkt, sizes = stochastic_ipm_system(7, nb, n_x, n_eq, n_in, n_fs, same_pattern=same); print("build s", time.perf_counter() - t0, "block rows", sizes[0], "coupling", sizes[-1])
rng = np.random.default_rng(0); rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
s = B200SchurComplementLinearSolver(options=opts)
t0 = time.perf_counter(); s.do_symbolic_factorization(kkt); torch.cuda.synchronize(); print("symbolic s", time.perf_counter() - t0)
print(s.backend.plan_stats(0), "factor GB", s.backend.factor_bytes() / 1e9)
for rep in range(3):
    try:
        t0 = time.perf_counter(); st = s.do_numeric_factorization(kkt, raise_on_error=False).status; torch.cuda.synchronize(); t1 = time.perf_counter()
    except RuntimeError as e:
        sts = [s.backend.plan_stats(b) for b in range(nb)]
        print("FAILED", e, [(x["failed"] % 10, (x["failed"] // 10) % 100, (x["failed"] // 1000) % 100000, x["failed"] // 100000000) for x in sts])
        sys.exit(0)
    if st.value != 0: print("status", st); break
    ine = s.get_inertia(); x = s.do_back_solve(rhs); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep", rep, "numeric ms", (t1 - t0) * 1e3, "solve ms", (t2 - t1) * 1e3, st, ine)
print([ (s.backend.plan_stats(b)["delayed_to_root"], s.backend.plan_stats(b)["fell_back_dense"]) for b in range(min(nb, 4))])
print({k: (round(v["ms"] / 3, 3), v["launches"] // 3) for k, v in s.backend.profile().items()})
K = sym_full(kkt).tocsr(); b = rhs.flatten()
print("rel residual", np.linalg.norm(K @ x.flatten() - b) / np.linalg.norm(b), "device estimate", s.last_residual, "refine steps", s.refine_steps)
s.max_refine = 0; x0 = s.do_back_solve(rhs)
print("rel residual without refinement", np.linalg.norm(K @ x0.flatten() - b) / np.linalg.norm(b))
n_prim = nb * (n_x + n_in) + n_fs
print("expected inertia if convex", (n_prim, sum(sizes) - n_prim, 0))

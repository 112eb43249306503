"""Config-5 share of one GPU at N = 8 (16 blocks x 20 000 rows x 2 000 coupling): per-class device times of a step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 2
extra = dict(kv.split("=") for kv in sys.argv[3:])
m = EstimationModel(nb, 2000, 4, 2000)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver(options={"profile": 1, "overlap_groups": groups, **{k: float(v) for k, v in extra.items()}})
s.do_symbolic_factorization(kkt)
for _ in range(2):
    s.do_numeric_factorization(kkt); x = s.do_back_solve(rhs)
s.backend.profile()
reps = 5
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    s.do_numeric_factorization(kkt); x = s.do_back_solve(rhs)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps * 1e3
p = s.backend.profile()
print(f"blocks {nb} groups {groups} {extra}: step {dt:.2f} ms  " + "  ".join(f"{k} {v['ms']/reps:.2f}" for k, v in p.items()), flush=True)
print("inertia ok", s.get_inertia() == m.expected_inertia(), "residual", s.last_residual, "max_err", m.check_result(x))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    x = s.do_back_solve(rhs)
torch.cuda.synchronize(); print(f"back_solve alone {(time.perf_counter() - t0) / reps * 1e3:.2f} ms")

"""Family-P systems (config-3 / config-4 shapes) with option overrides: per-class device times of a step.
usage: family_probe.py {stoch|dyn} <blocks> [key=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_families import dynamic_ipm_system, stochastic_ipm_system
from tests.helpers import block_vector
from parapint_b200 import B200SchurComplementLinearSolver
kind, nb = sys.argv[1], int(sys.argv[2])
extra = {k: float(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
if kind == "stoch":
    kkt, sizes = stochastic_ipm_system(7, nb, 10000, 8000, 1000, 200, same_pattern=True)
else:
    kkt, sizes = dynamic_ipm_system(9, nb, 5000, 4800, 100, 50, same_pattern=True)
rhs = block_vector(np.random.default_rng(11).standard_normal(sum(sizes)), sizes)
s = B200SchurComplementLinearSolver(options={"profile": 1, **extra})
s.do_symbolic_factorization(kkt)
for _ in range(2):
    s.do_numeric_factorization(kkt); x = s.do_back_solve(rhs)
s.backend.profile()
reps = 4
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    s.do_numeric_factorization(kkt); x = s.do_back_solve(rhs)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps * 1e3
p = s.backend.profile()
print(f"{kind} {nb} {extra}: step {dt:.2f} ms  " + "  ".join(f"{k} {v['ms']/reps:.2f}" for k, v in p.items()), " residual %.2e" % s.last_residual, s.backend.plan_stats(0)["delayed_to_root"], flush=True)

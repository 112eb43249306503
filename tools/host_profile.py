import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cProfile, pstats
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
m = EstimationModel(64, 150, 6, 50)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver()
s.do_symbolic_factorization(kkt)
def step():
    s.do_numeric_factorization(kkt); s.get_inertia(); return s.do_back_solve(rhs)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

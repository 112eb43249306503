"""Small driver for ncu: config-2 workload, one symbolic, `reps` x (numeric + inertia + solve)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = EstimationModel(blocks, 150, 6, 50)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver()
s.do_symbolic_factorization(kkt)
for _ in range(reps):
    assert s.do_numeric_factorization(kkt).status.value == 0
    assert s.get_inertia() == m.expected_inertia()
    x = s.do_back_solve(rhs)
print("max_err", m.check_result(x), "launches", s.backend.kernel_launches())

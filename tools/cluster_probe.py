"""Subtree kernels with 1/2/4/8 CTAs per block at several block counts (config-2 blocks): per-class device times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
for nb in (8, 16, 32, 64):
    m = EstimationModel(nb, 150, 6, 50)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    for c in (1, 2, 4, 8):
        if nb * c > 148:
            continue
        s = B200SchurComplementLinearSolver(options={"profile": 1, "subtree_cluster": c})
        s.do_symbolic_factorization(kkt)
        for _ in range(5):
            s.do_numeric_factorization(kkt); s.do_back_solve(rhs)
        s.backend.profile()
        reps = 50
        torch.cuda.synchronize()
        for _ in range(reps):
            s.do_numeric_factorization(kkt); s.do_back_solve(rhs)
        torch.cuda.synchronize()
        p = s.backend.profile()
        print(f"blocks {nb:3d} cluster {c}: subtree {p['subtree']['ms']/reps*1e3:7.1f}  forward {p['forward']['ms']/reps*1e3:7.1f}  backward {p['backward']['ms']/reps*1e3:7.1f} us", flush=True)
        assert s.get_inertia() == m.expected_inertia()

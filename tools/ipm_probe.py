import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.helpers import stochastic_ipm_system, block_vector
from parapint_b200 import B200SchurComplementLinearSolver
from oracle.schur_oracle import sym_full, dense_inertia
for opts in [{}, {"sparse_dslot": 16, "sparse_dmax": 48}, {"pivot_threshold": 1e-6}]:
  for args in [(0, 4, 300, 240, 30, 10), (1, 3, 600, 500, 20, 25), (2, 6, 220, 100, 60, 8)]:
    kkt, sizes = stochastic_ipm_system(*args)
    rng = np.random.default_rng(args[0]); rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
    o = dict(opts); o["no_fallback"] = 1
    s = B200SchurComplementLinearSolver(options=o)
    s.do_symbolic_factorization(kkt)
    try:
        r = s.do_numeric_factorization(kkt, raise_on_error=False)
        print(opts, args, r.status, [s.backend.plan_stats(b)["delayed_to_root"] for b in range(args[1])])
    except RuntimeError as e:
        st = [s.backend.plan_stats(b) for b in range(args[1])]
        print(opts, args, "FAILED", [(x["failed"] % 10, (x["failed"] // 10) % 100, (x["failed"] // 1000) % 100000, x["failed"] // 100000000) for x in st], {k: st[0][k] for k in ("supernodes", "root_cols", "max_front")})

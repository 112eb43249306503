"""Per-block accuracy of the config-4-shaped solve against SuperLU on the same blocks (host), to separate solver
error from the conditioning floor eps*|K||x|/|b| of the generated systems."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sp, scipy.sparse.linalg as spl
from tests.helpers import stochastic_ipm_system, block_vector
from parapint_b200 import B200SchurComplementLinearSolver
nb = int(sys.argv[1]); scale = float(sys.argv[2])
n_x, n_eq, n_in, n_fs = int(10000 * scale), int(8000 * scale), int(1000 * scale), int(200 * min(1.0, scale * 2))
kkt, sizes = stochastic_ipm_system(7, nb, n_x, n_eq, n_in, n_fs, same_pattern=True)
rng = np.random.default_rng(0); rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
s = B200SchurComplementLinearSolver()
s.do_symbolic_factorization(kkt); s.do_numeric_factorization(kkt); x = s.do_back_solve(rhs)
print("device residual", s.last_residual, "steps", s.refine_steps)
xc = np.asarray(x.get_block(nb)); bn = np.linalg.norm(rhs.flatten())
rows = []
for i in range(nb):
    K = kkt.get_block(i, i).tocsr(); A = kkt.get_block(nb, i).tocsr()
    xi = np.asarray(x.get_block(i)); bi = np.asarray(rhs.get_block(i))
    rows.append((np.linalg.norm(K @ xi + A.T @ xc - bi) / bn, np.linalg.norm(xi), i))
rows.sort(reverse=True)
print("worst blocks (res/|b|, |x_i|, i):", rows[:4]); print("best:", rows[-2:])
for res, nx, i in rows[:2] + rows[-1:]:
    K = kkt.get_block(i, i).tocsc(); A = kkt.get_block(nb, i).tocsr()
    bi = np.asarray(rhs.get_block(i)) - A.T @ xc
    t0 = time.perf_counter(); lu = spl.splu(K); y = lu.solve(bi); dt = time.perf_counter() - t0
    xi = np.asarray(x.get_block(i))
    print(f"block {i}: gpu res {np.linalg.norm(K @ xi - bi) / bn:.3e}  superlu res {np.linalg.norm(K @ y - bi) / bn:.3e}"
          f"  |x_gpu| {np.linalg.norm(xi):.3e} |x_lu| {np.linalg.norm(y):.3e} rel diff {np.linalg.norm(xi - y) / np.linalg.norm(y):.3e}  ({dt:.1f}s)")

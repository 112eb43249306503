import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import dynamic_ipm_system, block_vector
from oracle.schur_oracle import SchurOracle, sym_full
from parapint_b200 import B200SchurComplementLinearSolver, native, structure
seed, N, n_x, n_eq, n_in, n_s = 0, 12, 60, 30, 5, 4
kkt, sizes = dynamic_ipm_system(seed, N, n_x, n_eq, n_in, n_s)
rng = np.random.default_rng(seed); rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
o = SchurOracle(); o.symbolic(kkt); o.numeric(kkt); xr = o.solve(rhs)
S_ref = o.S - kkt.get_block(N, N).toarray()      # contribution only
st = structure.analyse(kkt)
res = {}
for name, opt in (("sparse", {"coupling_min_sparse": 16, "defer_status": 0, "auto_residual": 0}), ("dense", {"coupling_min_sparse": 10**6, "defer_status": 0, "auto_residual": 0})):
    s = B200SchurComplementLinearSolver(options=opt, max_refine=0)
    s.do_symbolic_factorization(kkt); print(name, s.do_numeric_factorization(kkt, raise_on_error=False).status, s.backend.coupling_stats())
    sch = s.backend.schur.cpu().numpy()
    x = s.do_back_solve(rhs)
    xc = np.asarray(x.get_block(N)).ravel()
    print(name, "inertia", s.get_inertia(), "xc err", np.linalg.norm(xc - np.asarray(xr.get_block(N)).ravel()) / np.linalg.norm(np.asarray(xr.get_block(N)).ravel()))
    res[name] = (sch, xc, s)
cl_ptr = st.border_ptr; cl_rows = st.border_rows
Q = kkt.get_block(N, N).tocoo(); keep = Q.row >= Q.col
plan = native.coupling_plan(st.m_c, cl_ptr, cl_rows, Q.row[keep], Q.col[keep], min_mc=16)
colidx = np.repeat(np.arange(st.m_c), np.diff(plan["colptr"]))
sp_vals = res["sparse"][0][: plan["nnz"]]
dn = res["dense"][0][: st.m_c ** 2].reshape(st.m_c, st.m_c).T
print("pattern nnz", plan["nnz"], "schur_size", res["sparse"][2].backend.schur_size)
d1 = dn[plan["rowidx"], colidx]
print("S_local sparse vs dense path: max abs diff", np.abs(sp_vals - d1).max(), "max", np.abs(d1).max())
print("S dense path vs oracle:", np.abs(np.tril(dn) - np.tril(S_ref)).max())
mask = np.zeros((st.m_c, st.m_c), bool); mask[plan["rowidx"], colidx] = True
print("oracle S entries outside pattern:", np.abs(np.tril(S_ref)[~mask]).max())

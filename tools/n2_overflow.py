"""2-rank check of the deferred-status flow: one rank's sparse path overflows (no delayed-pivot capacity), every rank
sees it in the reduced tail, the local phase is repeated densely; result compared with a dense solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from tests.helpers import stochastic_ipm_system, block_vector
from oracle.schur_oracle import sym_full, dense_inertia
from parapint_b200 import B200SchurComplementLinearSolver
from parapint_b200.comm import Communicator
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
comm = Communicator()
kkt, sizes = stochastic_ipm_system(3, 4, 300, 240, 30, 10)
rng = np.random.default_rng(3); rhs = block_vector(rng.standard_normal(sum(sizes)), sizes)
for opts in ({}, {"sparse_dmax": 0}):
    s = B200SchurComplementLinearSolver(comm=comm, options=opts)
    s.do_symbolic_factorization(kkt)
    for rep in range(2):
        s.do_numeric_factorization(kkt)
        x = s.do_back_solve(rhs)
    K = sym_full(kkt).toarray()
    xs = np.zeros(sum(sizes)); off = np.concatenate(([0], np.cumsum(sizes)))
    for i in range(len(sizes)):
        b = x.get_block(i)
        if b is not None and (i % comm.size == rank or i == len(sizes) - 1):
            xs[off[i]:off[i + 1]] = np.asarray(b).ravel()
    t = torch.from_numpy(xs).cuda()
    # coupling block is replicated: keep rank 0's copy only
    if rank != 0: t[off[-2]:] = 0
    dist.all_reduce(t); xs = t.cpu().numpy()
    res = np.linalg.norm(K @ xs - rhs.flatten()) / np.linalg.norm(rhs.flatten())
    ine = s.get_inertia()
    if rank == 0:
        print(opts, "residual", res, "inertia", ine, "expected", dense_inertia(K, "ldl"), "stats", s.backend.plan_stats(0)["fell_back_dense"], "defer", s._defer)
        assert res < 1e-10 and tuple(ine) == tuple(dense_inertia(K, "ldl"))
dist.destroy_process_group()

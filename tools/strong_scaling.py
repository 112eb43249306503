"""Strong scaling of a fixed system (default BASELINE config 5: 128 blocks x 20 000 rows x 2 000 coupling) over the
ranks of one node: torchrun --nproc-per-node N tools/strong_scaling.py [n_blocks n_q y_mult n_theta].  Every rank
builds only its own blocks (round-robin), times numeric factorisation + inertia + back-solve through the plugin
(host buffers), and rank 0 prints the max over ranks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
from parapint_b200.comm import Communicator
args = [int(a) for a in sys.argv[1:5]] or [128, 2000, 4, 2000]
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
comm = Communicator()
m = EstimationModel(*args, local_blocks=[i for i in range(args[0]) if i % world == rank])
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver(comm=comm)
t0 = time.perf_counter(); s.do_symbolic_factorization(kkt); torch.cuda.synchronize(); sym = time.perf_counter() - t0
times = []
for rep in range(4):
    comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    s.do_numeric_factorization(kkt); t1 = time.perf_counter()
    ine = s.get_inertia(); x = s.do_back_solve(rhs); torch.cuda.synchronize(); t2 = time.perf_counter()
    times.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t = torch.tensor(times[1:], device="cuda", dtype=torch.float64).median(dim=0).values
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"system": args, "n_gpus": world, "numeric_ms": float(t[0]), "solve_ms": float(t[1]), "symbolic_s": sym,
                      "inertia_ok": tuple(ine) == tuple(m.expected_inertia()), "residual_estimate": s.last_residual}))
if world > 1:
    dist.destroy_process_group()

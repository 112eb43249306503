import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
from bench import build_model
from parapint_b200 import B200SchurComplementLinearSolver, Communicator
comm = Communicator(); model = build_model(world, rank); kkt, rhs = model.build_kkt(), model.build_rhs()
solver = B200SchurComplementLinearSolver(comm=comm); be = solver.backend
solver.do_symbolic_factorization(kkt); solver.do_numeric_factorization(kkt); solver.get_inertia(); solver.do_back_solve(rhs)
values_dev = be.values_pin.to(dev); rhs_dev = be.rhs_pin.to(dev); rhsc_dev = be.rhsc_pin.to(dev)
x_dev = torch.empty_like(rhs_dev); xc_dev = torch.empty_like(rhsc_dev)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return time.perf_counter()
for it in range(13):
    if it == 3: T.clear()
    t = time.perf_counter()
    code, s_local = be.numeric_local_device(values_dev); t = tick("numeric_local", t)
    comm.allreduce_sum_(s_local); t = tick("allreduce_S", t)
    be.numeric_coupling(s_local); t = tick("coupling", t)
    be.inertia_local(); be.inertia_coupling(); t = tick("inertia", t)
    be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_); t = tick("solve", t)
if rank == 0:
    print({k: round(v / 10 * 1e3, 3) for k, v in T.items()})
# without per-piece syncs
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for it in range(10):
    code, s_local = be.numeric_local_device(values_dev)
    comm.allreduce_sum_(s_local)
    be.numeric_coupling(s_local)
    be.inertia_local(); be.inertia_coupling()
    be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_)
torch.cuda.synchronize(); dist.barrier()
if rank == 0: print("dev step ms", (time.perf_counter() - t0) / 10 * 1e3)
dist.destroy_process_group()

"""Where does the config-2 device step go?  Variants of bench.py's dev_step: profile events on / off, L2 flush on / off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver, Communicator
m = EstimationModel(64, 150, 6, 50); kkt, rhs = m.build_kkt(), m.build_rhs()
dev = torch.device("cuda", 0); comm = Communicator()
flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)
small = torch.empty(20 * 1024 * 1024, dtype=torch.float64, device=dev)
for prof in (1, 0):
    opts = {"profile": prof}
    for a in sys.argv[1:]:
        k, v = a.split("="); opts[k] = float(v)
    s = B200SchurComplementLinearSolver(comm=comm, options=opts); be = s.backend
    s.do_symbolic_factorization(kkt); s.do_numeric_factorization(kkt); s.get_inertia(); s.do_back_solve(rhs)
    values_dev, rhs_dev, rhsc_dev = be.values_pin.to(dev), be.rhs_pin.to(dev), be.rhsc_pin.to(dev)
    x_dev, xc_dev = torch.empty_like(rhs_dev), torch.empty_like(rhsc_dev)
    def dev_step(fl):
        if fl is not None: fl.fill_(1.0)
        code, s_local = be.numeric_local_device(values_dev)
        code2 = be.numeric_coupling(s_local)
        loc = be.inertia_local(); cpl = be.inertia_coupling()
        be.solve_device(rhs_dev, rhsc_dev, x_dev, xc_dev, reduce=comm.allreduce_sum_)
    for name, fl in (("flush 256 MB", flush), ("flush 160 MB", small), ("no flush", None)):
        for _ in range(5): dev_step(fl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(200): dev_step(fl)
        e1.record(); torch.cuda.synchronize()
        print(f"profile {prof} {name}: {e0.elapsed_time(e1) / 200:.4f} ms device, {(time.perf_counter() - t0) * 5:.4f} ms wall")
    if prof:
        p = be.profile(reset=True); print({k: round(v['ms'] / 615, 4) for k, v in p.items()})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fl in (flush, small):
    e0.record()
    for _ in range(50): fl.fill_(1.0)
    e1.record(); torch.cuda.synchronize(); print("fill alone", fl.numel() * 8 >> 20, "MB:", e0.elapsed_time(e1) / 50, "ms")

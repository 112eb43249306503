import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from parapint_b200 import native
native.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libparapint_b200_cstrace.so")
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = EstimationModel(nb, 2000, 4, 2000)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver(max_refine=0, options={"overlap_groups": 1})
s.do_symbolic_factorization(kkt); s.do_numeric_factorization(kkt)
lib = native.load()
buf = (C.c_longlong * 24)()
lib.pp_debug_pc_trace.argtypes = [C.c_void_p, C.c_int]
torch.cuda.synchronize(); lib.pp_debug_pc_trace(buf, 1)
s.do_numeric_factorization(kkt); torch.cuda.synchronize(); lib.pp_debug_pc_trace(buf, 1)
t = np.array(buf[:], dtype=np.int64)
cols = t[8]
us = t[:8] / 1965.0
names = ["loop top", "wrow+sync", "sweep 1", "argmax 1", "second column (sweep+argmax)", "sync", "interchange", "scale+sync"]
print("columns (block 0 root + coupling front):", cols, " second-column evaluations:", t[9], " interchanges:", t[10], " 2x2:", t[11])
for nme, v in zip(names, us):
    print(f"  {nme:32s} {v:9.1f} us total  {v / max(cols,1):6.2f} us/column")
print("  total %.1f us = %.2f us/column" % (us.sum(), us.sum() / max(cols, 1)))

panels = max(int(t[23]), 1)
sp = t[12:18] / 1965.0
for nme, v in zip(["load diagonal block", "phase A (block)", "publish + cluster sync", "phase B (rows)", "reduce + decide", "commit"], sp):
    print(f"  spec {nme:28s} {v:9.1f} us total  {v / panels:7.2f} us/panel")
print("  panels", panels, " speculative total %.2f us/panel" % (sp.sum() / panels))

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
s = B200SchurComplementLinearSolver(max_refine=0)
s.do_symbolic_factorization(kkt); s.do_numeric_factorization(kkt)
lib = native.load()
buf = (C.c_longlong * 8)()
lib.pp_debug_cs_trace.argtypes = [C.c_void_p, C.c_int]
x = s.do_back_solve(rhs); torch.cuda.synchronize(); lib.pp_debug_cs_trace(buf, 1)
x = s.do_back_solve(rhs); torch.cuda.synchronize(); lib.pp_debug_cs_trace(buf, 1)
t = np.array(buf[:], dtype=np.int64) / 1965.0
print("forward sweeps of block 0 (root, then coupling front), us: loop-top %.1f owner %.1f cluster-sync %.1f update %.1f block-sync %.1f total %.1f" % (t[0], t[1], t[2], t[3], t[4], t[:5].sum()))

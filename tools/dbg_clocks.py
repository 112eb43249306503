import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver, native
m = EstimationModel(64, 150, 6, 50)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver()
s.do_symbolic_factorization(kkt)
lib = native.load(); lib.pp_debug_clocks.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_longlong * 64)()
for rep in range(3):
    s.do_numeric_factorization(kkt)
    lib.pp_debug_clocks(buf)
v = list(buf)
print("phase cycles (block 0, CTA fronts): head+ndin %d | zero+fid %d | originals %d | stage-fetch %d | apply %d | factor %d | store %d" % (v[0], v[1], v[2], v[3], v[4], v[5], v[6]))
print("fronts %d children %d cols %d S %d ne %d ; tiny-pass cycles %d big-pass cycles %d" % (v[20], v[21], v[22], v[23], v[24], v[30], v[31]))

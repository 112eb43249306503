"""Per-level cycle stamps of the subtree factor / forward kernels (block 0), built with -DPP_TRACE:
   nvcc ... -DPP_TRACE -o tools/libparapint_b200_trace.so parapint_b200/csrc/capi.cu"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from parapint_b200 import native
native.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libparapint_b200_trace.so")
from oracle.kkt_generator import EstimationModel
from parapint_b200 import B200SchurComplementLinearSolver
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = EstimationModel(blocks, 150, 6, 50)
kkt, rhs = m.build_kkt(), m.build_rhs()
s = B200SchurComplementLinearSolver()
s.do_symbolic_factorization(kkt)
lib = native.load()
for _ in range(3):
    torch.cuda.synchronize(); lib.pp_debug_trace_reset()
    s.do_numeric_factorization(kkt); torch.cuda.synchronize()
    if os.environ.get("TRACE_SOLVE"): lib.pp_debug_trace_reset()
    x = s.do_back_solve(rhs)
torch.cuda.synchronize()
buf = (C.c_longlong * 2048)()
lib.pp_debug_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.pp_debug_trace(buf, 2048) == 0
t = np.array(buf[:], dtype=np.int64)
mhz = 1965.0
nl = s.backend.plan_stats(0).get("levels", None)
for base, name in ((0, "factor"), (1024, "forward")):
    tt = t[base:base + (256 if base == 0 else 1024)]
    nlev = int(np.max(np.nonzero(tt)[0]) // 4) if np.any(tt) else 0
    print(name, "levels", nlev, "total us", (tt[4 * nlev] - tt[0]) / mhz if name == "factor" else (tt[4 * nlev - 1] - tt[0]) / mhz)
    for l in range(nlev):
        a, b, c, d = tt[4 * l:4 * l + 4]
        print(f"  level {l:2d}: tiny {(b - a) / mhz:7.2f}  med {(c - b) / mhz:7.2f}  big {(d - c) / mhz:7.2f} us")
    if name == "factor":
        print(f"  root extend-add: {(tt[4 * nlev] - tt[4 * nlev - 1]) / mhz:7.2f} us")

ph = t[256:1016]; ph = ph[ph != 0]
if os.environ.get("TRACE_SOLVE"):
    names = ["start", "L loaded + inv", "v init", "children counted", "children fetched", "children applied", "tri solve", "out stored", "-"]
else:
    names = ["start", "children counted", "zero+ids", "orig entries", "children fetched", "children applied", "big children", "factor", "stored"]
print("process_front<128> phases (group 0 of block 0), us since previous stamp:")
prev = None
for v in ph[:200]:
    cyc, p = int(v) >> 4, int(v) & 15
    if p == 0: print("  ---")
    if prev is not None and p != 0: print(f"    {names[p]:18s} {(cyc - prev) / mhz:7.2f}")
    prev = cyc

print("factor_front<128> on block 0 / group 0 over 3 factorizations: no-pivot, 1x1, 2x2 steps:", t[2040:2043], "candidate columns examined:", t[2043])

sm = t[1999:2016]
print("front_small_kernel, block 0: roots  load %.2f  sync %.2f  pivot block %.2f  border rows %.2f  trailing %.2f  store %.2f us" % tuple((sm[i + 1] - sm[i]) / mhz for i in range(6)))
print("front_small_kernel, coupling: pivot block %.2f  store %.2f us" % ((sm[11] - sm[10]) / mhz, (sm[14] - sm[11]) / mhz))

"""Farmer LP (BASELINE config 1): trajectory of the SciPy-leaf convention (|lambda| <= 1e-8 counts as zero) against the
B200 solver with an absolute zero-pivot tolerance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ipm import StochasticInterface, farmer_scenarios, ip_solve
from oracle.schur_oracle import OraclePlugin
from parapint_b200 import B200SchurComplementLinearSolver
def run(solver):
    scen, fs, _ = farmer_scenarios()
    return ip_solve(StochasticInterface(scen, fs), solver)
ref = run(OraclePlugin())
print("scipy-leaf convention:", ref["iterations"], ref["objective"], [r for r in ref["reg"] if r[1] > 0][:6])
for tol in (0.0, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6):
    out = run(B200SchurComplementLinearSolver(options={"pivot_tol": tol}))
    same = [(r[1], r[2]) for r in out["reg"]] == [(r[1], r[2]) for r in ref["reg"]]
    print("pivot_tol", tol, out["status"], out["iterations"], out["objective"], "same reg decisions:", same, [r for r in out["reg"] if r[1] > 0][:4])

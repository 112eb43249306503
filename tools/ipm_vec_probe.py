"""Achieved HBM bandwidth of the interior-point vector kernels (N3) on vectors larger than L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parapint_b200.ipm_vectors import IpmKernels
k = IpmKernels()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000     # config 4: 1024 scenarios x 10 000 primals ~ 1e7
g = torch.Generator(device=k.device); g.manual_seed(0)
lb = -torch.rand(n, dtype=torch.float64, device=k.device, generator=g) - 0.1
ub = torch.rand(n, dtype=torch.float64, device=k.device, generator=g) + 0.1
x = torch.zeros(n, dtype=torch.float64, device=k.device)
dx = torch.randn(n, dtype=torch.float64, device=k.device, generator=g)
zl = torch.rand(n, dtype=torch.float64, device=k.device, generator=g)
zu = torch.rand(n, dtype=torch.float64, device=k.device, generator=g)
out = torch.ones(8, dtype=torch.float64, device=k.device)
alpha = torch.tensor([1e-3, 1e-3, 1.0], dtype=torch.float64, device=k.device)
res = {}
def timed(name, fn, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[name] = {"ms": ms, "GB/s": nbytes / ms * 1e-6, "bytes": nbytes}
timed("fraction_to_boundary (6 vectors read)", lambda: k.fraction_to_boundary(out[:2], 0.99, 1e-2, x, dx, lb, ub, zl, zu), 48 * n)
timed("complementarity (5 vectors read)", lambda: k.complementarity(out[:6], 1e-2, x, lb, ub, zl, zu), 40 * n)
timed("max_abs (2 vectors read)", lambda: k.max_abs(out[:2], x, dx), 16 * n)
timed("step (6 read, 3 written)", lambda: k.step(alpha, 1e-2, x, dx, lb, ub, zl, zu), 72 * n)
timed("axpy (2 read, 1 written)", lambda: k.axpy(alpha, 1, x, dx), 24 * n)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6543.1
for v in res.values(): v["frac_of_measured_hbm_peak"] = v["GB/s"] / peak
print(json.dumps({"n": n, "hbm_peak_gbs": peak, "kernels": res}, indent=1))

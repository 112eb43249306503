"""cuBLAS DGEMM / HBM copy probe (roofline denominators not present in MEASURED_PEAKS.json)."""
import json, torch
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
# sustained
n = 8192
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40):
    c = a @ b
e1.record(); torch.cuda.synchronize()
res["dgemm_8192_sustained_tflops"] = 40 * 2 * n**3 / e0.elapsed_time(e1) * 1e-9
x = torch.empty(1 << 28, dtype=torch.float64, device=dev); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["hbm_copy_gbs"] = 2 * x.numel() * 8 / best * 1e-6
print(json.dumps(res))

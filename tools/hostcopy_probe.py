import sys, time, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parapint_b200 import native
print("cpus", os.cpu_count())
N=129; arrs=[np.random.rand(4081) for _ in range(N)]; offs=np.cumsum([0]+[a.size for a in arrs])[:-1]
out=np.zeros(sum(a.size for a in arrs))
big=np.random.rand(out.size)
t=time.perf_counter()
for _ in range(100): np.copyto(out, big)
print("single memcpy 4.2MB ms", (time.perf_counter()-t)/100*1e3)
for th in (1,2,4,8):
    cp=native.HostCopier(th)
    for _ in range(5): cp.copy(arrs, offs, out)
    t=time.perf_counter()
    for _ in range(100): cp.copy(arrs, offs, out)
    print(th, "threads ms", (time.perf_counter()-t)/100*1e3)

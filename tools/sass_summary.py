"""SASS instruction counts per kernel of the built library (cuobjdump -sass): which kernels hold DMMA / cp.async /
cluster barriers.  python tools/sass_summary.py > profiles/sass_summary_rNN.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "parapint_b200", "csrc", "libparapint_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ("DMMA", "LDGSTS", "LDG", "LD.", "STG", "ST.", "LDS", "STS", "BAR", "DFMA", "SHFL", "MATCH", "REDUX", "UCGABAR")
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cnt[cur]["total"] += 1
        for key in KEYS:
            if op.startswith(key):
                cnt[cur][key] += 1
print("SASS instruction counts per kernel of libparapint_b200.so (cuobjdump -sass; sm_100a).  DMMA = FP64 tensor-core MMA")
print("(mma.sync.m8n8k4.f64), LDGSTS = cp.async, UCGABAR = cluster barrier, MATCH = match.any, REDUX = redux.sync,")
print("LD. / ST. = generic-address loads / stores.")
print(f"{'kernel':46s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
for k, c in cnt.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0].replace("ppb::", "").replace("void ", "")
    print(f"{name[:46]:46s} {c['total']:6d} " + " ".join(f"{c[key]:7d}" for key in KEYS))

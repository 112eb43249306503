"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")
print(f"{'total':60s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")

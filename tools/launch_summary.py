"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (durations are cold-cache and
serialised under the profiler: use the shares, not the absolute values)."""
import csv, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
body = rows[hdr + 1:]
if len(sys.argv) > 3:  # launch_summary.py file.csv <kernel substring> <K>: only from the K-th last launch of that kernel on
    idx = [i for i, r in enumerate(body) if sys.argv[2] in r[kn]]
    body = body[idx[-int(sys.argv[3])]:] if len(idx) >= int(sys.argv[3]) else body
agg = OrderedDict()
for r in body:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    if r[mu] in ("ns", "nsecond"):
        v /= 1e3
    elif r[mu] in ("ms", "msecond"):
        v *= 1e3
    name = r[kn].split("(")[0][:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} n={n:4d} total={t:10.1f} us avg={t / n:8.1f} us share={100 * t / tot:5.1f}%")
print(f"total us {tot:.1f} launches {sum(a[0] for a in agg.values())}")

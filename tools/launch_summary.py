"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= v:
        continue
    name = re.sub(r"\(.*", "", r[k]).replace("void <unnamed>::", "").replace("void ", "")
    try:
        t = float(r[v].replace(",", ""))
    except ValueError:
        continue
    t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u], 1e-6)
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {n[:80]} | {c} | {t:.3f} | {100 * t / tot:.1f} % | {1e3 * t / c:.1f} |")

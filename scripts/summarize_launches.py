"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "")
    t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f}% |")
print(f"| all | {sum(v[0] for v in agg.values())} | {tot:.0f} | 100% |")

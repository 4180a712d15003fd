"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/summarise_launches.py launches.csv[.gz] > summary.md"""
import csv
import gzip
import sys

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    rows = list(csv.reader(l for l in f if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, total = {}, 0.0
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
    k = r[ik].split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v; total += v
print("| kernel | launches | total us | avg us | share |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, n, t, t / n, 100 * t / total))
print("| **all** | %d | %.1f | | |" % (sum(a[0] for a in agg.values()), total))

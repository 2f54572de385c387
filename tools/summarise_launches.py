"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total / average / share.
    python tools/summarise_launches.py gpurun_out/launches.csv [first_id last_id]   (ids select one training step)"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {n: i for i, n in enumerate(hdr)}
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = OrderedDict()
for r in rows:
    i = int(r[ix["ID"]])
    if not (lo <= i <= hi):
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("icl::", "")
    if "gemm" in name or "rec_" in name:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]].replace("icl::", "")).strip()
    name += " grid=" + r[ix["Grid Size"]].replace(" ", "") if "--grid" in sys.argv else ""
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ix["Metric Value"]]) / 1e3
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (n[:90], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
print("\nTotal: %d launches, %.2f ms serialised kernel time." % (sum(a[0] for a in agg.values()), tot / 1e3))

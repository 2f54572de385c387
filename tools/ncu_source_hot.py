"""Summarise the SASS page of one kernel of an ncu report: the instructions that collect the most warp-stall samples, with the
dominant stall reason of each, and the totals per stall reason.
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv ; python tools/ncu_source_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
name = rows[0][1] if rows[0][0] == "Kernel Name" else "?"
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[hdr]
ci = {c: i for i, c in enumerate(cols)}
stall_cols = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
end = next((i for i in range(hdr + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))      # first launch only
body = [r for r in rows[hdr + 1:end] if len(r) == len(cols)]
samp = lambda r: int(r[ci["# Samples"]] or 0)
total = sum(samp(r) for r in body)
print("kernel: %s\nSASS instructions: %d, warp-stall samples: %d\n" % (name, len(body), total))
print("| stall reason | samples | share |\n|---|---:|---:|")
tot = {c: sum(int(r[ci[c]] or 0) for r in body) for c in stall_cols}
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("| %s | %d | %.1f%% |" % (c, v, 100.0 * v / max(1, total)))
print("\n| # | SASS index | instruction | samples | share | dominant stall |\n|---:|---:|---|---:|---:|---|")
order = sorted(range(len(body)), key=lambda i: -samp(body[i]))[:top]
for n, i in enumerate(order):
    r = body[i]
    dom = max(stall_cols, key=lambda c: int(r[ci[c]] or 0))
    print("| %d | %d | `%s` | %d | %.1f%% | %s |" % (n + 1, i, " ".join(r[ci["Source"]].split()), samp(r), 100.0 * samp(r) / max(1, total), dom))

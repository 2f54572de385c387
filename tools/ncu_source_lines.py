"""Aggregate the warp-stall samples of an ncu report by CUDA source line.
  ncu -i rep.ncu-rep --page source --print-source sass,cuda --csv --kernel-name regex:NAME > sc.csv ; python tools/ncu_source_lines.py sc.csv [top [function-substring]]
(needs a capture made with --import-source on and a library built with -lineinfo; only the first launch in the file is read)"""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else ""        # substring of the function name when the file holds several kernels
out, fpath, cols, func0, func, seen = [], None, None, None, None, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        func = r[1]
        if want in func:
            func0 = func0 or func
        continue
    if r[0] == "Line No":
        cols = r; continue
    if cols and func == func0 and len(r) == len(cols) and r[0].strip().isdigit():
        if (fpath, int(r[0])) not in seen:                      # a second launch of the same kernel repeats every line
            seen.add((fpath, int(r[0])))
            out.append((fpath, int(r[0]), r[1], r))
ci = {}
for i, c in enumerate(cols):
    ci.setdefault(c, i)
stall = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
S = lambda r: int(r[ci["# Samples"]]) if r[ci["# Samples"]].isdigit() else 0
total = sum(S(r) for _, _, _, r in out)
print("kernel: %s\nwarp-stall samples: %d over %d source lines\n" % (func0, total, len(out)))
print("| # | file:line | samples | share | top stall reasons | source |\n|---:|---|---:|---:|---|---|")
for n, (f, ln, src, r) in enumerate(sorted(out, key=lambda t: -S(t[3]))[:top]):
    dom = [c for c in sorted(stall, key=lambda c: -(int(r[ci[c]]) if r[ci[c]].isdigit() else 0))[:2] if r[ci[c]].isdigit() and int(r[ci[c]]) > 0]
    print("| %d | %s:%d | %d | %.1f%% | %s | `%s` |" % (n + 1, f, ln, S(r), 100.0 * S(r) / max(1, total),
                                                     ", ".join("%s %s" % (c[6:], r[ci[c]]) for c in dom), " ".join(src.split())[:120].replace("|", "\\|")))

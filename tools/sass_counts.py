"""SASS instruction counts per kernel of the built library -> profiles/<tag>_sass_counts.md (runs here: cuobjdump needs no GPU).
    python tools/sass_counts.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "imagecaptionlearn_py_b200", "libicl_b200.so")
COLS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKPF", r"\bUBLKPF"),
        ("UCGABAR", r"\bUCGABAR"), ("SYNCS", r"\bSYNCS"), ("MUFU", r"\bMUFU"), ("RED/ATOM", r"\b(RED|ATOM|ATOMG|REDG)\b")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, cur, order = collections.OrderedDict(), None, iter(names)
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(order)
            cur = re.sub(r"^void ", "", cur)
            cur = re.sub(r"\(.*$", "", cur).replace("icl::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        counts[cur]["instr"] += 1
        for col, pat in COLS:
            if re.match(pat.replace(r"\b", ""), m.group(1)) if col != "RED/ATOM" else re.match(r"(RED|ATOM|ATOMG|REDG)(\.|$)", m.group(1)):
                counts[cur][col] += 1
    tot = collections.Counter()
    rows = []
    for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["instr"]):
        tot.update(c)
        rows.append("| `%s` | %d | %s |" % (k, c["instr"], " | ".join(str(c[col]) for col, _ in COLS)))
    out = os.path.join(ROOT, "profiles", "%s_sass_counts.md" % tag)
    with open(out, "w") as f:
        f.write("# SASS evidence (%s): `cuobjdump -sass imagecaptionlearn_py_b200/libicl_b200.so`, instruction counts per kernel\n\n" % tag)
        f.write("Built by `__graft_entry__.build()` (`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3`); regenerate with "
                "`python tools/sass_counts.py %s`. `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` / `UTMASTG` = TMA "
                "tensor loads / stores (cp.async.bulk.tensor), `UBLKPF` = cp.async.bulk.prefetch.L2, `UCGABAR` = cluster barrier, `SYNCS` = "
                "mbarrier operations, `MUFU` = special-function unit.\n\n" % tag)
        f.write("Library totals: %d kernels, %s.\n\n" % (len(counts), ", ".join("%d %s" % (tot[col], col) for col, _ in COLS)))
        f.write("| kernel | SASS instr. | %s |\n|---|---:|%s\n" % (" | ".join(c for c, _ in COLS), "---:|" * len(COLS)))
        f.write("\n".join(rows) + "\n")
    print(out, "kernels", len(counts), dict((col, tot[col]) for col, _ in COLS))


if __name__ == "__main__":
    main()

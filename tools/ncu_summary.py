"""Per-kernel summary of an `ncu --set full` report (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/r1f_full.ncu-rep [out.json]
duration, DRAM bytes read+written, DRAM throughput %, tensor-pipe active %, registers, achieved occupancy."""
import csv, io, json, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
want = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor_pct2": "sm__inst_executed_pipe_tensor.sum", "regs": "launch__registers_per_thread",
        "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
def scale(v, u):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
                "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(u, 1)
out = []
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("icl::", "").replace("void ", "")
    if "gemm" in name:
        name = re.sub(r"\(CUtensorMap.*", "", r[ix["Kernel Name"]].replace("icl::", "").replace("void ", ""))
    e = dict(id=int(r[ix["ID"]]), kernel=name.strip(), grid=r[ix["Grid Size"]], block=r[ix["Block Size"]])
    for k, m in want.items():
        if m in ix:
            e[k] = scale(r[ix[m]], units[ix[m]])
    if "dram_rd" in e:
        e["dram_bytes"] = e["dram_rd"] + e["dram_wr"]
        e["dram_gbs"] = e["dram_bytes"] / (e["dur_us"] * 1e-6) / 1e9
    out.append(e)
print("| id | kernel | grid | us | DRAM MB (rd+wr) | DRAM GB/s | dram % | L2 % | tensor pipe % | regs |")
print("|---:|---|---|---:|---:|---:|---:|---:|---:|---:|")
for e in out:
    print("| %d | `%s` | %s | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d |" % (
        e["id"], e["kernel"][:48], e["grid"].replace(" ", ""), e["dur_us"], e.get("dram_bytes", 0) / 1e6, e.get("dram_gbs", 0),
        e.get("dram_pct", float("nan")), e.get("l2_pct", float("nan")), e.get("tensor_pct", float("nan")), int(e.get("regs", 0))))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)

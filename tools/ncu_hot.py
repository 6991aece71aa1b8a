"""Hot instructions of one ncu --set full --import-source on capture: stall samples by opcode and the top instructions.
usage: python tools/ncu_hot.py report.ncu-rep [ntop]"""
import collections, csv, subprocess, sys
rep, ntop = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, v = raw[0], raw[-1]
g = lambda k: next((v[i] for i, x in enumerate(h) if x == k), "?")
print(g("Kernel Name")[:90]); print("grid", g("Grid Size"), "time us", g("gpu__time_duration.sum"), "regs", g("launch__registers_per_thread"),
      "fmaheavy%", g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), "issue%", g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
      "dram rd/wr", g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), "warps%", g("sm__warps_active.avg.pct_of_peak_sustained_active"))
st = sorted(((float(v[i]), x) for i, x in enumerate(h) if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("_per_issue_active.ratio")), reverse=True)
print("stalls:", ", ".join("%s %.2f" % (x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), f) for f, x in st[:7]))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]; ix = {x: i for i, x in enumerate(hdr)}; data = rows[2:]
S = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
cat = collections.Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    cat[".".join(op.split(".")[:2]) if "WIDE" in op else op.split(".")[0]] += S(r)
print("samples", tot, "|", ", ".join("%s %.1f%%" % (k, 100 * s / tot) for k, s in cat.most_common(9)))
for r in sorted(data, key=lambda r: -S(r))[:ntop]:
    print("  %5d %4.1f%%  %-70s long_sb %s short_sb %s wait %s lg %s" % (S(r), 100 * S(r) / tot, r[ix["Source"]][:70], r[ix["stall_long_sb"]], r[ix["stall_short_sb"]], r[ix["stall_wait"]], r[ix["stall_lg"]]))

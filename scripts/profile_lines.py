#!/usr/bin/env python3
"""Summarise an `ncu --set full --import-source on` capture of one kernel into two tracked files under profiles/:
selected raw metrics (CSV) and the top source lines by stall samples / warp instructions.

    scripts/profile_lines.py gpurun_out/X.ncu-rep profiles/<name> "<title line>" """
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_active.avg.per_cycle_active", "launch__", "issue_stalled",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "Kernel Name")
with open(out + "_ncu_raw.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keep) and "not_issued" not in h:
            w.writerow([h, u, v])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h2 = None; agg = {}
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h2 = r; continue
    if h2 is None or not r[0].isdigit() or r[2] != "-": continue
    i_s = h2.index("# Samples"); i_i = h2.index("Instructions Executed"); i_t = h2.index("Thread Instructions Executed")
    a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
    a[0] += int(r[i_s] or 0); a[1] += int(r[i_i] or 0); a[2] += int(r[i_t] or 0)
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
with open(out + "_lines.txt", "w") as f:
    f.write(title + "\n")
    f.write("total warp instructions %d, stall samples %d; top 30 source lines by samples\n" % (ti, ts))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        f.write("%-20s %5d samples %5.1f%%  instructions %5.1f%%  active threads %4.1f  %s\n" % (k[0], k[1], 100 * a[0] / max(1, ts), 100 * a[1] / max(1, ti), a[2] / max(1, a[1]), a[3].strip()[:100]))
print(open(out + "_lines.txt").read()[:1500])

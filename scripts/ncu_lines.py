#!/usr/bin/env python3
"""Aggregate an ncu source page (cuda,sass CSV) per source line: instructions executed, samples.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv; ncu_lines.py both.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if not r[0].isdigit(): continue
    if r[2] != "-": continue  # sass rows carry an address; keep the per-line summary rows
    try:
        i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed")
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, r[1]])
    a[0] += int(r[i_samp] or 0); a[1] += int(r[i_inst] or 0); a[2] += int(r[i_thr] or 0)
tot_i = sum(a[1] for a in agg.values()); tot_s = sum(a[0] for a in agg.values())
print("total warp-instr %d, samples %d" % (tot_i, tot_s))
print("%-24s %5s %7s %7s %5s  %s" % ("file", "line", "inst%", "samp%", "thr", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-24s %5d %6.2f%% %6.2f%% %5.1f  %s" % (key[0], key[1], 100.0 * a[1] / tot_i, 100.0 * a[0] / max(1, tot_s), a[2] / max(1, a[1]), a[3].strip()[:100]))

# optional: regions "name:lo-hi,name:lo-hi" as argv[3] (lines of the main .cu file)
if len(sys.argv) > 3:
    print()
    main = sys.argv[4] if len(sys.argv) > 4 else "inflate_parallel.cu"
    regs = []
    for part in sys.argv[3].split(","):
        nm, rng = part.split(":"); lo, hi = rng.split("-"); regs.append((nm, int(lo), int(hi)))
    sums = {nm: [0, 0, 0] for nm, _, _ in regs}; other = [0, 0, 0]
    for (f, ln), a in agg.items():
        hit = None
        if f == main:
            for nm, lo, hi in regs:
                if lo <= ln <= hi: hit = nm; break
        t = sums[hit] if hit else other
        t[0] += a[0]; t[1] += a[1]; t[2] += a[2]
    for nm, _, _ in regs + [("other", 0, 0)]:
        t = sums.get(nm, other)
        print("%-12s inst %6.2f%%  samples %6.2f%%  avg threads %5.1f" % (nm, 100.0 * t[1] / tot_i, 100.0 * t[0] / max(1, tot_s), t[2] / max(1, t[1])))

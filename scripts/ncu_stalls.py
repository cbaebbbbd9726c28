#!/usr/bin/env python3
"""Per source line: which stall reason the warp samples of an ncu source page (cuda,sass CSV) sit in.
Usage: ncu_stalls.py both.csv <stall column, e.g. stall_barrier> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
col = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
hdr = None; cur = None; agg = {}; tot_all = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or r[2] != "-": continue
    i = hdr.index(col); i_s = hdr.index("# Samples")
    v = int(r[i] or 0); tot_all += int(r[i_s] or 0)
    if v:
        a = agg.setdefault((cur, int(r[0])), [0, r[1]]); a[0] += v
tot = sum(a[0] for a in agg.values())
print("%s: %d samples of %d (%.1f%%)" % (col, tot, tot_all, 100.0 * tot / max(1, tot_all)))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %5d %6.2f%%  %s" % (f, l, 100.0 * a[0] / max(1, tot), a[1].strip()[:110]))

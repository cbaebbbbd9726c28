#!/usr/bin/env python3
"""Turn a `ncu --set full` capture brought back in gpurun_out/ into the tracked summaries under profiles/.

    scripts/make_profile_summary.py <tag> <streams> [round]

reads  gpurun_out/<tag>_prof.ncu-rep, gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_bench.json
writes profiles/r<round>_<tag>_ncu_raw.csv      (selected raw metrics of the captured launch)
       profiles/r<round>_<tag>_launches.csv     (gpu__time_duration per launch, the recipe's launch list)
       profiles/r<round>_<tag>_bench.json       (the bench line of the same build)
       profiles/r<round>_<tag>_lines.txt        (instructions / stall samples per source line and per phase)
       profiles/inflate_parallel_ncu_summary.json (what bench.py reads for roofline.traffic)
"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, streams = sys.argv[1], int(sys.argv[2])
rnd = sys.argv[3] if len(sys.argv) > 3 else "01"
rep = os.path.join(ROOT, "gpurun_out", tag + "_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__", "issue_stalled",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "Kernel Name", "dram__cycles_active")
out = os.path.join(ROOT, "profiles", "r%s_%s_ncu_raw.csv" % (rnd, tag))
vals_by = {}
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keep) and "not_issued" not in h:
            w.writerow([h, u, v])
            vals_by[h] = (u, v)

def num(name):
    u, v = vals_by[name]
    x = float(v.replace(",", ""))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return x * scale

summary = {"kernel": "dcmp::inflate_parallel_kernel", "streams": streams, "capture": os.path.basename(rep),
           "duration_s_under_ncu": num("gpu__time_duration.sum"), "dram_bytes_read": num("dram__bytes_read.sum"),
           "dram_bytes_write": num("dram__bytes_write.sum"), "warp_instructions": num("smsp__inst_executed.sum"),
           "ipc_per_sm": num("sm__inst_executed.avg.per_cycle_elapsed"),
           "threads_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "registers_per_thread": num("launch__registers_per_thread")}
json.dump(summary, open(os.path.join(ROOT, "profiles", "inflate_parallel_ncu_summary.json"), "w"), indent=1)
for ext, dst in (("_launches.csv", "_launches.csv"), ("_bench.json", "_bench.json")):
    src = os.path.join(ROOT, "gpurun_out", tag + ext)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(ROOT, "profiles", "r%s_%s%s" % (rnd, tag, dst)))
print(json.dumps(summary))

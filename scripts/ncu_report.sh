#!/bin/bash
# usage: scripts/ncu_report.sh gpurun_out/X_prof.ncu-rep "regions"
REP=$1
ncu -i $REP --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/both.csv
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_active.avg.per_cycle_active','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_bytes.sum']
for i,h in enumerate(hdr):
    if h in want: print('%-70s %s %s'%(h, rows[1][i], vals[i]))
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.3: print('%-70s %s'%(h.replace('smsp__average_warps_issue_stalled_','stall/issue: ').replace('_per_issue_active.ratio',''), vals[i]))
"
python scripts/ncu_lines.py /tmp/both.csv ${3:-25} "$2"

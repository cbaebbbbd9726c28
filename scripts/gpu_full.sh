#!/bin/bash
# Full GPU pass: every -m gpu test, the headline bench, the other configurations, sanity of the launch list.
# Usage: scripts/gpu_full.sh <tag>
TAG=${1:-full}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
timeout 900 python bench_extra.py --what c1,c3,c5 2>&1 | grep "^{" | tee $OUT/${TAG}_extra.json | cut -c1-330
timeout 600 python bench_extra.py --what c5 --c5-blocks 1024 2>&1 | grep "^{" | tee -a $OUT/${TAG}_extra.json | cut -c1-330
for lv in 1 4 9; do timeout 600 python bench_extra.py --what c3 --c3-level $lv 2>&1 | grep "^{" | tee -a $OUT/${TAG}_extra.json | cut -c1-260; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_c3_launches.csv \
    python bench_extra.py --what c3 --c3-mib 256 --steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1
grep -c deflate $OUT/${TAG}_c3_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_cta -s 1 -c 1 -f -o $OUT/${TAG}_deflate_prof \
    python bench_extra.py --what c3 --c3-mib 256 --steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_deflate_prof.ncu-rep

#!/bin/bash
# quick deflate iteration: C3 at the given levels (+ candidate counts) + one ncu capture.
# Usage: scripts/gpu_deflate_quick.sh <tag> "<levels>" "<ways to try at level 6>" [mib]
TAG=${1:-dq}; LV=${2:-6}; WAYS=${3:-}; MIB=${4:-1024}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_deflate.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest.txt
for lv in $LV; do
  timeout 600 python bench_extra.py --what c3 --c3-level $lv --c3-mib $MIB 2>&1 | tail -1 | cut -c1-260 | tee -a $OUT/${TAG}_c3.json
done
for w in $WAYS; do
  echo "ways $w"; DCMP_DC_WAYS=$w timeout 600 python bench_extra.py --what c3 --c3-level 6 --c3-mib $MIB 2>&1 | tail -1 | cut -c1-260 | tee -a $OUT/${TAG}_c3_ways.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_cta -s 1 -c 1 -f -o $OUT/${TAG}_prof \
    python bench_extra.py --what c3 --c3-mib 256 --steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep

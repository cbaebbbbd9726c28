#!/bin/bash
# Deflate iteration on the GPU box: parity tests, C3 at levels 1/6/9 (new kernel) and level 6 on the round-1 kernel,
# per-kernel times, one ncu capture of the CTA-per-chunk kernel.   Usage: scripts/gpu_deflate.sh <tag> [tests]
TAG=${1:-def}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "none" ]; then
  timeout 1500 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_gzip.py tests/test_gpu_streaming.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
fi
for lv in 6 1 9; do
  timeout 600 python bench_extra.py --what c3 --c3-level $lv 2>&1 | tail -2 | tee -a $OUT/${TAG}_c3.json
done
DCMP_DEF_KERNEL=warp timeout 600 python bench_extra.py --what c3 2>&1 | tail -1 | tee -a $OUT/${TAG}_c3_warp.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_c3_launches.csv \
    python bench_extra.py --what c3 --c3-mib 256 --steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1
grep -c deflate $OUT/${TAG}_c3_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_cta -s 1 -c 1 -f -o $OUT/${TAG}_prof \
    python bench_extra.py --what c3 --c3-mib 256 --steps 1 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep

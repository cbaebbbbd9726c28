#!/bin/bash
# One tuning iteration on the GPU box: inflate parity tests, bench, variants, ncu launch list + full capture.
# Usage: scripts/gpu_iter.sh <tag> [tests]   (tests: "all" | "inflate" | "none")
TAG=${1:-it}
TESTS=${2:-inflate}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt
if [ "$TESTS" = "all" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
elif [ "$TESTS" = "inflate" ]; then
  timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_gzip.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
fi
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
echo "== variants"
for f in variants_*.so; do
  [ -f "$f" ] || continue
  echo "== $f"
  DCMP_LIB_PATH=$PWD/$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f GB/s  ms/step %.3f  e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee $OUT/${TAG}_variants.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
grep -c inflate $OUT/${TAG}_launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_parallel -s 3 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep

#!/bin/bash
# C5 (LZO1X blocks of 1 MiB) on several tuning builds of the library (variants_*.so at the repo root)
N=${1:-512}
mkdir -p gpurun_out
for f in variants_*.so; do
  echo "== $f"
  DCMP_LIB_PATH=$PWD/$f timeout 600 python bench_extra.py --what c5 --c5-blocks $N 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   %.1f GB/s  %.2f ms'%(d['value'], d['ms']))"
done 2>&1 | tee gpurun_out/lzo_variants.txt

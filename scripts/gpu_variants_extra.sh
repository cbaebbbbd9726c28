#!/bin/bash
# bench_extra.py configurations ($1, e.g. c3 or c5) for several tuning builds of the library (variants_*.so at the repo root)
mkdir -p gpurun_out
for f in variants_*.so; do
  echo "== $f"
  DCMP_LIB_PATH=$PWD/$f timeout 600 python bench_extra.py --what ${1:-c3} 2>&1 | grep -E '^\{' | cut -c1-260
done 2>&1 | tee gpurun_out/variants_extra.txt

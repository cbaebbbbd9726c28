#!/bin/bash
# bench several tuning builds of the library (variants_*.so at the repo root)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1   # generate + cache the workload
for f in variants_*.so; do
  echo "== $f"
  DCMP_LIB_PATH=$PWD/$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   value %.1f GB/s  ms/step %.3f  e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:200])
"
done 2>&1 | tee gpurun_out/variants.txt

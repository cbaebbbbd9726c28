#!/bin/bash
# e2e (host buffers) throughput of the headline bench for several pipeline slice sizes
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for mb in 64 48 32 24 16; do
  echo -n "slice ${mb} MB: "
  DCMP_SLICE_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f  e2e %.2f GB/s'%(d['value'],d['e2e']['value']))"
done

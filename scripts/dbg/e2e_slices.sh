#!/bin/bash
# e2e (host buffers) throughput of the headline bench: doubling slices with several first-slice sizes, and round 1's even slices
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f  e2e %.2f GB/s  ceiling %.1f'%(d['value'],d['e2e']['value'],d['e2e']['pcie_ceiling']))"; }
for mb in 2 4 8 16 32; do echo -n "doubling, first slice ${mb} MB: "; DCMP_SLICE_MB=$mb run; done
echo -n "even slices of 48 MB (round 1): "; DCMP_SLICE_EVEN=1 DCMP_SLICE_MB=48 run
echo -n "even slices of 24 MB: "; DCMP_SLICE_EVEN=1 DCMP_SLICE_MB=24 run

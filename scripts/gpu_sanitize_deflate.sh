#!/bin/bash
# compute-sanitizer memcheck + racecheck over the CTA-per-chunk compressor (and the small-stream kernel beside it).
mkdir -p gpurun_out
cat > /tmp/sand.py <<'PY'
import sys; sys.path.insert(0,'.')
import zlib, numpy as np, workloads
from decompress_b200 import _lib, batch
from tests import cases
eng=_lib.Engine(0)
txt=workloads.text(300000,3)
noise=np.random.default_rng(1).integers(0,256,70000,dtype=np.uint8).tobytes()
edge=cases.deflate_window_edge_stream()
ins=[txt, b"", txt[:100], bytes(140000), b"abc", edge, noise, txt[:65537], (b"abcdefghij"*7+b"\n")*2000, edge[5:]]
for lv in (0,1,6,9):
    d=batch.deflate_streams(ins,level=lv,fmt=batch.FMT_ZLIB,engine=eng)
    assert all(r[0]==0 and zlib.decompress(r[2])==x for r,x in zip(d,ins)), lv
d=batch.deflate_streams(ins,level=6,fmt=_lib.FMT_GZIP,engine=eng); assert all(r[0]==0 for r in d)
small=[txt[k*37:k*37+20+k*11] for k in range(100)]   # many small streams: the warp-per-chunk kernel, outputs packed for one transfer
d=batch.deflate_streams(small,level=6,fmt=batch.FMT_ZLIB,engine=eng); assert all(r[0]==0 and zlib.decompress(r[2])==x for r,x in zip(d,small))
print("sanitizer workload ok")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/sand.py 2>&1 | tail -6 | tee gpurun_out/r2_memcheck_deflate.log
cat > /tmp/sand2.py <<'PY'
import sys; sys.path.insert(0,'.')
import zlib, workloads
from decompress_b200 import _lib, batch
eng=_lib.Engine(0)
txt=workloads.text(140000,3)
d=batch.deflate_streams([txt, bytes(70000)],level=6,fmt=batch.FMT_ZLIB,engine=eng); assert zlib.decompress(d[0][2])==txt
print("racecheck workload ok")
PY
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/sand2.py > gpurun_out/r2_racecheck_deflate.log 2>&1
grep -E "Race reported|hazard|RACECHECK SUMMARY|workload ok" gpurun_out/r2_racecheck_deflate.log | sed -e 's/.*deflate.cu:/deflate.cu:/' | sort | uniq -c | sort -rn | head -20

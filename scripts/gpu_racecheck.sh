#!/bin/bash
# compute-sanitizer racecheck + memcheck over small inflate / LZO / deflate workloads through the C ABI
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, zlib
sys.path.insert(0, ".")
import workloads
from decompress_b200 import _lib, batch, lzo
eng = _lib.Engine(0)
comps, plains = workloads.c2_streams(6, workers=1)
res = batch.inflate_streams(comps, [65536] * 6, fmt=batch.FMT_ZLIB, engine=eng)
assert all(r[0] == 0 and r[3] == p for r, p in zip(res, plains))
txt = workloads.text(300000, 9)
co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
multi = co.compress(txt) + co.flush()
res = batch.inflate_streams([multi, zlib.compress(bytes(100000))[2:-4]], [len(txt), 100000], engine=eng)
assert res[0][3] == txt and res[1][0] == 0
blocks, raws = workloads.c5_blocks(2)
res = lzo.uncompress_batch(blocks + [workloads.lzo1x_1_compress(bytes(200000)), workloads.lzo1x_1_compress(b"ab" * 90000)],
                           [len(r) for r in raws] + [200000, 180000], engine=eng)
assert all(r[0] == 0 for r in res)
res = batch.deflate_streams([txt, txt[:1000], b""], level=6, fmt=batch.FMT_ZLIB, engine=eng)
assert all(zlib.decompress(c) == x for (st, w, c), x in zip(res, [txt, txt[:1000], b""]))
print("sanitizer workload ok")
PY
for tool in racecheck memcheck; do
  echo "== compute-sanitizer $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 40 python /tmp/san_small.py > gpurun_out/r2_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok" gpurun_out/r2_$tool.log
  grep -E "hazard detected" gpurun_out/r2_$tool.log | sed -E 's/.*(Race reported|hazard detected)/\1/' | sort | uniq -c | head -12
  grep -E "^=+ +.* in .*\.(cu|cuh):[0-9]+" gpurun_out/r2_$tool.log | sed -E 's/.* in //' | sort | uniq -c | sort -rn | head -12
done

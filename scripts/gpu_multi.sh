#!/bin/bash
# 2-GPU bench through torchrun (the driver's launch line), smoke(), and a compute-sanitizer pass on a small batch
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench x2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -2 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-400
echo "== reference arm x2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
echo "== compute-sanitizer memcheck (inflate + deflate + lzo, small)"
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
import workloads, zlib, numpy as np
from decompress_b200 import _lib, batch, lzo
from tests import cases
eng=_lib.Engine(0)
comps,plains=workloads.c2_streams(8,workers=1)
r=batch.inflate_streams(comps,[65536]*8,fmt=batch.FMT_ZLIB,engine=eng); assert all(x[0]==0 and x[3]==p for x,p in zip(r,plains))
vs=cases.valid_streams(); r=batch.inflate_streams([c for _,c,_ in vs],[len(p) for _,_,p in vs],engine=eng); assert all(x[3]==p for x,(_,_,p) in zip(r,vs))
ms=cases.mutated_streams(60); batch.inflate_streams([s for _,s in ms],[1<<17]*len(ms),engine=eng)
txt=workloads.text(200000,3); d=batch.deflate_streams([txt,b"",txt[:100],bytes(70000)],level=6,fmt=batch.FMT_ZLIB,engine=eng); assert zlib.decompress(d[0][2])==txt
lz=[workloads.lzo1x_1_compress(txt[:50000]), workloads.lzo1x_1_compress(bytes(3000))]; r=lzo.uncompress_batch(lz,[50000,3000],engine=eng); assert r[0][3]==txt[:50000]
print("sanitizer workload ok")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py 2>&1 | tail -6
echo "== compute-sanitizer racecheck (inflate, small)"
cat > /tmp/san2.py <<'PY'
import sys; sys.path.insert(0,'.')
import workloads
from decompress_b200 import _lib, batch
eng=_lib.Engine(0)
comps,plains=workloads.c2_streams(2,workers=1)
r=batch.inflate_streams(comps,[65536]*2,fmt=batch.FMT_ZLIB,engine=eng); assert all(x[0]==0 and x[3]==p for x,p in zip(r,plains))
print("racecheck workload ok")
PY
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san2.py 2>&1 | tail -8

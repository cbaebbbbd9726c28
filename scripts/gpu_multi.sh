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
import workloads, zlib, gzip, numpy as np
from decompress_b200 import _lib, batch, lzo
from tests import cases
eng=_lib.Engine(0)
comps,plains=workloads.c2_streams(8,workers=1)
r=batch.inflate_streams(comps,[65536]*8,fmt=batch.FMT_ZLIB,engine=eng); assert all(x[0]==0 and x[3]==p for x,p in zip(r,plains))
vs=cases.valid_streams(); r=batch.inflate_streams([c for _,c,_ in vs],[len(p) for _,_,p in vs],engine=eng); assert all(x[3]==p for x,(_,_,p) in zip(r,vs))
ms=cases.mutated_streams(60); batch.inflate_streams([s for _,s in ms],[1<<17]*len(ms),engine=eng)
txt=workloads.text(200000,3); d=batch.deflate_streams([txt,b"",txt[:100],bytes(70000)],level=6,fmt=batch.FMT_ZLIB,engine=eng); assert zlib.decompress(d[0][2])==txt
noise=np.random.default_rng(1).integers(0,256,40000,dtype=np.uint8).tobytes()
lzp=[txt[:150000], bytes(30000), noise, txt[:1000]+noise+txt[1000:2000], b"ab"*20000]
lz=[workloads.lzo1x_1_compress(p) for p in lzp]; r=lzo.uncompress_batch(lz,[len(p) for p in lzp],engine=eng); assert all(x[0]==0 and x[3]==p for x,p in zip(r,lzp))
r=lzo.uncompress_batch([c[:len(c)//2] for c in lz]+[c[:7]+b"\xff"+c[8:] for c in lz],[1<<18]*10,mode=1,engine=eng)
co=zlib.compressobj(6,zlib.DEFLATED,-15,8); mb=co.compress(txt)+co.flush()   # several blocks
st0=zlib.compressobj(0,zlib.DEFLATED,-15); sb=st0.compress(noise)+st0.flush()  # stored blocks
r=batch.inflate_streams([mb,sb,mb[:len(mb)//2]],[len(txt),len(noise),len(txt)],engine=eng); assert r[0][3]==txt and r[1][3]==noise
gzs=[gzip.compress(txt[:k]) for k in (10,5000,100000)]; r=batch.inflate_streams(gzs,[10,5000,100000],fmt=2,engine=eng); assert all(x[0]==0 for x in r)
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
from decompress_b200 import lzo
t=workloads.text(60000,3); c=workloads.lzo1x_1_compress(t); r=lzo.uncompress_batch([c],[60000],engine=eng); assert r[0][3]==t
print("racecheck workload ok")
PY
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/racecheck.log 2>&1; grep -E "Race reported|hazards\]|RACECHECK SUMMARY|workload ok" gpurun_out/racecheck.log | sed -e 's/dcmp::[a-z_]*(.*)+0x[0-9a-f]* in //' | sort | uniq -c | sort -rn | head -30

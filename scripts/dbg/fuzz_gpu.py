"""Differential fuzz on the GPU (not part of the suite): random payloads / levels / strategies / flush patterns
through dcmp_inflate_batch vs system zlib (the reference's own differential oracle, fuzz/fuzz.ml:82-88), LZO blocks
vs minilzo, deflate round trips."""
import os, random, sys, zlib
sys.path.insert(0, '.')
import numpy as np
import workloads
from decompress_b200 import _lib, batch, lzo
eng = _lib.Engine(0)
rnd = random.Random(int(os.environ.get("SEED", "1")))
txt = workloads.text(1 << 20, 77)
def payload():
    k = rnd.randrange(6)
    n = rnd.choice([0, 1, 7, 100, 3000, 20000, 65536, 150000, 400000])
    if k == 0: return txt[rnd.randrange(1000):][:n]
    if k == 1: return bytes(rnd.getrandbits(8) for _ in range(min(n, 40000)))
    if k == 2: return bytes(n)
    if k == 3: return bytes(rnd.choice(b"ab") for _ in range(min(n, 60000)))
    if k == 4: return (txt[:rnd.randrange(1, 300)] * (n // 50 + 1))[:n]
    return b"".join(txt[rnd.randrange(900000):][:rnd.randrange(1, 400)] + bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 40))) for _ in range(n // 200 + 1))[:n]
streams, plains = [], []
for _ in range(600):
    p = payload()
    co = zlib.compressobj(rnd.randrange(0, 10), zlib.DEFLATED, -15, rnd.randrange(1, 10), rnd.choice([0, 1, 2, 3, 4]))
    s = b""
    if rnd.random() < 0.3 and len(p) > 10:
        cut = rnd.randrange(len(p))
        s = co.compress(p[:cut]) + co.flush(rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])) + co.compress(p[cut:])
    else:
        s = co.compress(p)
    streams.append(s + co.flush()); plains.append(p)
res = batch.inflate_streams(streams, [len(p) for p in plains], engine=eng)
bad = [k for k, (r, p, s) in enumerate(zip(res, plains, streams)) if not (r[0] == 0 and r[1] == len(s) and r[3] == p)]
print("inflate raw: %d streams, bad %s" % (len(streams), bad[:5]))
zs = [zlib.compress(p, rnd.randrange(1, 10)) for p in plains[:300]]
res = batch.inflate_streams(zs, [len(p) + rnd.randrange(0, 3) for p in plains[:300]], fmt=batch.FMT_ZLIB, engine=eng)
bad = [k for k, (r, p, s) in enumerate(zip(res, plains, zs)) if not (r[0] == 0 and r[1] == len(s) and r[3] == p)]
print("inflate zlib: %d streams, bad %s" % (len(zs), bad[:5]))
lz = [workloads.lzo1x_1_compress(p) for p in plains[:300]]
res = lzo.uncompress_batch(lz, [len(p) for p in plains[:300]], engine=eng)
bad = [k for k, (r, p, s) in enumerate(zip(res, plains, lz)) if not (r[0] == 0 and r[1] == len(s) and r[3] == p)]
print("lzo: %d blocks, bad %s" % (len(lz), bad[:5]))
for level in (1, 6, 9):
    d = batch.deflate_streams(plains[:200], level=level, fmt=batch.FMT_ZLIB, engine=eng)
    bad = [k for k, (r, p) in enumerate(zip(d, plains)) if not (r[0] == 0 and zlib.decompress(r[2]) == p)]
    print("deflate level %d: %d inputs, bad %s" % (level, 200, bad[:5]))

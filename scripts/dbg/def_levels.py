"""deflate throughput / ratio per level on 256 MiB of the C3 text (device-resident, CUDA events)"""
import sys, time, zlib; sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
import workloads
from decompress_b200 import _lib, batch
eng = _lib.Engine(0)
t = workloads.text(64 << 20, 0)
for level in (1, 6, 9):
    t0 = time.time(); r = batch.deflate_streams([t], level=level, fmt=batch.FMT_ZLIB, engine=eng)[0]; dt = time.time() - t0
    assert r[0] == 0 and zlib.decompress(r[2]) == t
    print("level %d: ratio %.3f  (host-call wall %.1f ms for 64 MiB incl. PCIe)" % (level, len(t) / r[1], dt * 1e3))

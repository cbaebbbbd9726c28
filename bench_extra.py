#!/usr/bin/env python3
"""bench_extra.py — the other BASELINE.json configurations (not the driver's headline line; that is bench.py).

    python bench_extra.py [--what c1,c3,c5] [--c3-mib 1024] [--c5-blocks 256] [--steps 5]

One JSON line per configuration: device-resident throughput (CUDA events), algorithmic-bytes roofline fraction,
and a CPU reference beside it (the oracle port for inflate / LZO; system zlib level 6, the reference's own
yardstick in bench/b.ml, for deflate).  Results are checked (round trip / oracle) outside the timed region."""
import argparse
import ctypes as C
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
from bench import load_peaks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c1,c3,c4,c5")
    ap.add_argument("--c4-mib", type=int, default=512)
    ap.add_argument("--c3-mib", type=int, default=1024)
    ap.add_argument("--c3-level", type=int, default=6)
    ap.add_argument("--c5-blocks", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    from decompress_b200 import _lib, batch
    from tests import oracle

    dev = torch.device("cuda", 0)
    eng = _lib.Engine(0)
    peak, _ = load_peaks()
    L = _lib.lib()
    t = lambda a: torch.from_numpy(a).to(dev)
    P = lambda x: C.c_void_p(x.data_ptr())
    HP = lambda a: C.c_void_p(a.ctypes.data)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps * 1e-3

    what = args.what.split(",")
    if "c1" in what:
        # C1: maximal stored blocks (0xFFFF payload), batched
        n = 2048
        payload = np.random.default_rng(0).integers(0, 256, 65535, dtype=np.uint8).tobytes()
        blk = workloads.stored_block(payload)
        src, so, sl, dst, do, dc = batch.pack([blk] * n, [65535] * n)
        d = [t(x) for x in (src, so, sl, dst, do, dc)]
        cons, prod = torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev)
        fn = lambda: batch.inflate_device(batch.FMT_RAW, d[0], d[1], d[2], d[3], d[4], d[5], cons, prod, st, engine=eng)
        sec = timed(fn, args.steps)
        assert int(st.abs().sum()) == 0 and d[3][:65535].cpu().numpy().tobytes() == payload
        algo = 2.0 * 65535 * n
        print(json.dumps({"config": "C1: %d stored blocks of 65535 B (BTYPE=00), batched" % n, "metric": "inflate_uncompressed_GBps",
                          "value": round(65535 * n / sec / 1e9, 2), "ms": round(sec * 1e3, 3),
                          "roofline": {"achieved": round(algo / sec / 1e9, 1), "peak": peak, "frac": round(algo / sec / 1e9 / peak, 4)}}), flush=True)
    if "c3" in what:
        size = args.c3_mib << 20
        txt = workloads.text(size, 0)
        src = np.frombuffer(txt, dtype=np.uint8)
        cap = int(L.dcmp_deflate_bound(size, batch.FMT_ZLIB))
        d_src = t(src.copy())
        d_dst = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
        wr = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        z, sl_, dc_ = np.zeros(1, dtype=np.int64), np.array([size], dtype=np.int64), np.array([cap], dtype=np.int64)

        def fn():
            rc = L.dcmp_deflate_batch_device(eng.handle, args.c3_level, batch.FMT_ZLIB, 1, P(d_src), HP(z), HP(sl_), P(d_dst), HP(z), HP(dc_),
                                             P(wr), P(st), C.c_void_p(stream))
            _lib.check(rc, "dcmp_deflate_batch_device")
        sec = timed(fn, args.steps)
        w = int(wr.item())
        assert int(st.item()) == 0
        comp = d_dst[:w].cpu().numpy().tobytes()
        assert zlib.decompress(comp) == txt, "deflate output does not round-trip through system zlib"
        # round trip through the reference's Zl.Inf.Ns.inflate (the CPU oracle's restatement), at full size
        obuf = np.zeros(size + 16, dtype=np.uint8)
        oc, op = C.c_int64(0), C.c_int64(0)
        rc = oracle.lib().orc_zl_inflate_ns(comp, w, C.c_void_p(obuf.ctypes.data), size, C.byref(oc), C.byref(op))
        assert (rc, oc.value, op.value) == (0, w, size) and obuf[:size].tobytes() == txt, "oracle Inf round trip failed"
        del obuf
        t0 = time.time()
        zc = zlib.compress(txt[:64 << 20], 6)
        zsec = time.time() - t0
        algo = size + w
        print(json.dumps({"config": "C3: level-%d deflate of %d MiB synthetic text as one zlib stream (128 KiB blocks, 32 KiB window carried across them)" % (args.c3_level, args.c3_mib),
                          "metric": "deflate_input_GBps", "value": round(size / sec / 1e9, 2), "ms": round(sec * 1e3, 2),
                          "ratio": round(size / w, 3), "zlib6_ratio": round((64 << 20) / len(zc), 3),
                          "roofline": {"achieved": round(algo / sec / 1e9, 1), "peak": peak, "frac": round(algo / sec / 1e9 / peak, 4)},
                          "cpu_baseline": {"value": round((64 << 20) / zsec / 1e9, 4), "unit": "GB/s", "cores": 1, "kind": "system zlib level 6 (bench/b.ml's yardstick)",
                                           "sample": "first 64 MiB of the same text"}}), flush=True)
    if "c4" in what:
        # C4 at one GPU's share: concatenated gzip members (16..256 KiB payloads), boundaries from a sidecar index;
        # header parse, DEFLATE body and CRC-32 / ISIZE verification all run on the device inside the timed region.
        members = workloads.c4_members(args.c4_mib << 20, seed=0)
        blobs = [m for m, _ in members]
        plains = [p for _, p in members]
        n = len(blobs)
        src, so, sl, dst, do, dc = batch.pack(blobs, [len(p) for p in plains])
        d = [t(x) for x in (src, so, sl, dst, do, dc)]
        cons, prod = torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev)
        fn = lambda: batch.inflate_device(_lib.FMT_GZIP, d[0], d[1], d[2], d[3], d[4], d[5], cons, prod, st, engine=eng)
        sec = timed(fn, args.steps)
        assert int(st.abs().sum()) == 0 and cons.cpu().numpy().tolist() == [len(b) for b in blobs]
        out = d[3].cpu().numpy()
        for k in range(0, n, max(1, n // 50)):
            assert out[do[k]:do[k] + len(plains[k])].tobytes() == plains[k], k
        total = float(sum(len(p) for p in plains))
        algo = float(sl.sum()) + total
        print(json.dumps({"config": "C4 (one GPU's share): %d gzip members, %d MiB inflated, payloads 16..256 KiB, several DEFLATE blocks each; "
                                    "header, body, CRC-32 and ISIZE on the device" % (n, args.c4_mib),
                          "metric": "inflate_uncompressed_GBps", "value": round(total / sec / 1e9, 2), "ms": round(sec * 1e3, 3),
                          "roofline": {"achieved": round(algo / sec / 1e9, 1), "peak": peak, "frac": round(algo / sec / 1e9 / peak, 4)}}), flush=True)
    if "c5" in what:
        n = args.c5_blocks
        comps, plains = workloads.c5_blocks(n)
        src, so, sl, dst, do, dc = batch.pack(comps, [1 << 20] * n)
        d = [t(x) for x in (src, so, sl, dst, do, dc)]
        cons, prod = torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev)

        def fn():
            rc = L.dcmp_lzo_uncompress_batch_device(eng.handle, 0, n, P(d[0]), P(d[1]), P(d[2]), P(d[3]), P(d[4]), P(d[5]), P(cons), P(prod),
                                                    P(st), C.c_void_p(stream))
            _lib.check(rc, "dcmp_lzo_uncompress_batch_device")
        sec = timed(fn, args.steps)
        out = d[3].cpu().numpy()
        assert int(st.abs().sum()) == 0
        for k in (0, n // 2, n - 1):
            assert out[do[k]:do[k] + (1 << 20)].tobytes() == plains[k]
        t0 = time.time()
        for k in range(min(n, 32)):
            oracle.lzo_uncompress(comps[k], 1 << 20)
        csec = (time.time() - t0) / min(n, 32)
        algo = float(sl.sum()) + n * (1 << 20)
        print(json.dumps({"config": "C5: %d LZO1X blocks of 1 MiB (minilzo lzo1x_1_compress of synthetic text, ratio %.2f)" % (n, n * (1 << 20) / float(sl.sum())),
                          "metric": "lzo_uncompressed_GBps", "value": round(n * (1 << 20) / sec / 1e9, 2), "ms": round(sec * 1e3, 2),
                          "roofline": {"achieved": round(algo / sec / 1e9, 1), "peak": peak, "frac": round(algo / sec / 1e9 / peak, 4)},
                          "cpu_baseline": {"value": round((1 << 20) / csec / 1e9, 4), "unit": "GB/s", "cores": 1, "kind": "port",
                                           "sample": "oracle/lzo.c on %d of the same blocks, one core" % min(n, 32)}}), flush=True)


if __name__ == "__main__":
    main()

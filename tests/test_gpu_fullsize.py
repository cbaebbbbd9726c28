"""GPU parity at BASELINE.json's stated sizes ("close the parity gaps at the stated sizes"):

* C2 at all 4096 streams, every output byte against the CPU oracle (oracle/inflate_ns.c, one stream per thread);
* C3: level-6 deflate of 256 MiB of text as one zlib stream, inflated back by the ORACLE's `Zl.Inf.Ns.inflate`
  restatement (not only by system zlib), `consumed == written`;
* the reference's own 15 corpus files (test/corpus, committed as data under tests/golden/corpus): deflate at levels
  0..9 -> oracle Inf and zlib, and zlib-made streams -> GPU inflate (test/test_deflate.ml:20-120, test_ns.ml:1096-1187);
* the two-stage dispatch took the FAST kernels for every valid unit of C2 / C4 / C5 (dcmp_engine_stats: nothing redone
  by the general kernels)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import _lib, batch, lzo
from tests import oracle

pytestmark = pytest.mark.gpu
CORPUS = os.path.join(os.path.dirname(__file__), "golden", "corpus")
THREADS = min(32, os.cpu_count() or 1)


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


def _oracle_batch(fmt, src, so, sl, do, dc):
    dst = np.zeros(int(do[-1] + dc[-1]) + 16, dtype=np.uint8)
    _, st, prod = oracle.bench_batch(fmt, src, so, sl, dst, do, dc, THREADS)
    return dst, st, prod


def test_c2_all_4096_streams_every_byte_vs_oracle(eng):
    comps, plains = workloads.c2_streams(4096)
    src, so, sl, dst, do, dc = batch.pack(comps, [65536] * len(comps))
    before = eng.stats()
    cons, prod, st = batch.inflate_host(batch.FMT_ZLIB, src, so, sl, dst, do, dc, engine=eng)
    after = eng.stats()
    odst, ost, oprod = _oracle_batch(batch.FMT_ZLIB, src, so, sl, do, dc)
    assert (st == 0).all() and (ost == 0).all()
    assert (prod == 65536).all() and (oprod == 65536).all()
    assert cons.tolist() == [len(c) for c in comps]
    n = int(do[-1] + dc[-1])
    assert np.array_equal(dst[:n], odst[:n])                      # every byte of all 4096 streams
    assert dst[do[17]:do[17] + 65536].tobytes() == plains[17]     # (and the oracle agrees with the generator)
    assert after["inflate_streams"] - before["inflate_streams"] == 4096
    assert after["inflate_redone"] == before["inflate_redone"], "a valid C2 stream fell back to the general kernel"


def test_c3_256mib_level6_roundtrip_through_oracle_inf(eng):
    size = 256 << 20
    txt = workloads.text(size, 0)
    cap = int(_lib.lib().dcmp_deflate_bound(size, batch.FMT_ZLIB))
    (st, written, comp), = batch.deflate_streams([txt], [cap], level=6, fmt=batch.FMT_ZLIB, engine=eng)
    assert st == 0 and written == len(comp) and written < size // 2
    src = np.frombuffer(comp, dtype=np.uint8)
    z, sl, dc = np.zeros(1, dtype=np.int64), np.array([written], dtype=np.int64), np.array([size], dtype=np.int64)
    out = np.zeros(size + 16, dtype=np.uint8)
    L = oracle.lib()
    cons, prod = C.c_int64(0), C.c_int64(0)
    rc = L.orc_zl_inflate_ns(comp, written, C.c_void_p(out.ctypes.data), size, C.byref(cons), C.byref(prod))
    assert (rc, cons.value, prod.value) == (0, written, size)     # the reference's Zl.Inf.Ns.inflate: Ok (written, size)
    assert out[:size].tobytes() == txt
    assert zlib.decompress(comp) == txt                           # the reference's interop oracle (test/bin/simple.t)
    del out, src, z, sl, dc


def _corpus():
    return [(f, open(os.path.join(CORPUS, f), "rb").read()) for f in sorted(os.listdir(CORPUS))]


@pytest.mark.parametrize("level", list(range(10)))
def test_corpus_deflate_every_level_inflates_under_oracle_and_zlib(level, eng):
    files = _corpus()
    assert len(files) == 15
    ins = [d for _, d in files]
    res = batch.deflate_streams(ins, level=level, fmt=batch.FMT_ZLIB, engine=eng)
    comps = []
    for (name, data), (st, wr, comp) in zip(files, res):
        assert st == 0 and wr == len(comp), name
        rc, cons, prod, out = oracle.zl_inflate_ns(comp, len(data) + 8)
        assert (rc, cons, prod) == (0, wr, len(data)) and out == data, name
        assert zlib.decompress(comp) == data, name
        comps.append(comp)
    back = batch.inflate_streams(comps, [len(d) for d in ins], fmt=batch.FMT_ZLIB, engine=eng)
    for (name, data), (st, cons, prod, out) in zip(files, back):
        assert (st, prod) == (0, len(data)) and out == data, name


def test_corpus_zlib_streams_inflate_like_the_oracle(eng):
    files = _corpus()
    streams, plains = [], []
    for name, data in files:
        for level, mem in ((1, 8), (6, 8), (9, 9), (6, 1)):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, mem)
            streams.append(co.compress(data) + co.flush())
            plains.append(data)
    for caps in ([len(p) for p in plains], [len(p) - 1 for p in plains]):
        res = batch.inflate_streams(streams, caps, engine=eng)
        for k, (st, cons, prod, out) in enumerate(res):
            rc, ocons, oprod, oout = oracle.inflate_ns(streams[k], caps[k])
            assert (st, prod, out) == (rc, oprod, oout), (files[k // 4][0], k % 4)
            if rc == 0:
                assert cons == ocons and out == plains[k]


def test_valid_c4_members_and_c5_blocks_never_reach_the_general_kernels(eng):
    members = workloads.c4_members(96 << 20, seed=0)
    blobs, plains = [m for m, _ in members], [p for _, p in members]
    s0 = eng.stats()
    res = batch.inflate_streams(blobs, [len(p) for p in plains], fmt=_lib.FMT_GZIP, engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod) == (0, len(blobs[k]), len(plains[k])) and out == plains[k], k
    s1 = eng.stats()
    assert s1["inflate_streams"] - s0["inflate_streams"] == len(blobs) and s1["inflate_redone"] == s0["inflate_redone"]
    comps, raw = workloads.c5_blocks(48)
    res = lzo.uncompress_batch(comps, [len(p) for p in raw], engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod) == (0, len(comps[k]), len(raw[k])) and out == raw[k], k
    s2 = eng.stats()
    assert s2["lzo_blocks"] - s1["lzo_blocks"] == len(comps) and s2["lzo_redone"] == s1["lzo_redone"]

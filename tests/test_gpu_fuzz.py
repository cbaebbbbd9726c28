"""Differential fuzz through the C ABI, the way the reference fuzzes itself: random payloads, levels, strategies
and flush patterns compressed by system zlib must inflate identically (fuzz/fuzz.ml:82-88, 169-194), whatever
minilzo compresses must decode identically (fuzz/fuzz_lzo.ml:30-39), and whatever we deflate must inflate back
under zlib (fuzz/fuzz_ns.ml:49-65).  Fixed seeds."""
import random
import zlib

import pytest

import workloads
from decompress_b200 import _lib, batch, lzo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


def _payloads(seed, count):
    rnd = random.Random(seed)
    txt = workloads.text(1 << 20, 77)

    def one():
        k = rnd.randrange(6)
        n = rnd.choice([0, 1, 7, 100, 3000, 20000, 65536, 150000, 400000])
        if k == 0:
            return txt[rnd.randrange(1000):][:n]
        if k == 1:
            return bytes(rnd.getrandbits(8) for _ in range(min(n, 40000)))
        if k == 2:
            return bytes(n)
        if k == 3:
            return bytes(rnd.choice(b"ab") for _ in range(min(n, 60000)))
        if k == 4:
            return (txt[:rnd.randrange(1, 300)] * (n // 50 + 1))[:n]
        return b"".join(txt[rnd.randrange(900000):][:rnd.randrange(1, 400)] + bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 40)))
                        for _ in range(n // 200 + 1))[:n]

    return rnd, [one() for _ in range(count)]


@pytest.mark.parametrize("seed", [1, 2])
def test_zlib_streams_inflate_identically(seed, eng):
    rnd, plains = _payloads(seed, 400)
    streams = []
    for p in plains:
        co = zlib.compressobj(rnd.randrange(0, 10), zlib.DEFLATED, -15, rnd.randrange(1, 10), rnd.choice([0, 1, 2, 3, 4]))
        if rnd.random() < 0.3 and len(p) > 10:
            cut = rnd.randrange(len(p))
            s = co.compress(p[:cut]) + co.flush(rnd.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])) + co.compress(p[cut:])
        else:
            s = co.compress(p)
        streams.append(s + co.flush())
    res = batch.inflate_streams(streams, [len(p) for p in plains], engine=eng)
    for k, (r, p, s) in enumerate(zip(res, plains, streams)):
        assert r[:3] == (0, len(s), len(p)) and r[3] == p, k
    zs = [zlib.compress(p, rnd.randrange(1, 10)) for p in plains[:200]]
    res = batch.inflate_streams(zs, [len(p) + rnd.randrange(0, 3) for p in plains[:200]], fmt=batch.FMT_ZLIB, engine=eng)
    for k, (r, p, s) in enumerate(zip(res, plains, zs)):
        assert r[:3] == (0, len(s), len(p)) and r[3] == p, k


def test_minilzo_blocks_decode_identically(eng):
    _, plains = _payloads(3, 300)
    lz = [workloads.lzo1x_1_compress(p) for p in plains]
    res = lzo.uncompress_batch(lz, [len(p) for p in plains], engine=eng)
    for k, (r, p, s) in enumerate(zip(res, plains, lz)):
        assert r[:3] == (0, len(s), len(p)) and r[3] == p, k


@pytest.mark.parametrize("level", [1, 6, 9])
def test_deflate_inflates_back_under_zlib(level, eng):
    _, plains = _payloads(4, 200)
    d = batch.deflate_streams(plains, level=level, fmt=batch.FMT_ZLIB, engine=eng)
    for k, (r, p) in enumerate(zip(d, plains)):
        assert r[0] == 0 and zlib.decompress(r[2]) == p, k

"""GPU parity tests for the deflate path (dcmp_deflate_batch through the C ABI).

Deflate output is not unique and the reference's parse is sequential (lib/de.ml:4351-4410), so parity is the
reference's own acceptance rule for compressors (fuzz/fuzz.ml:169-194, test/test_ns.ml:1096-1187): what we emit must
inflate back to the input byte for byte under the reference's Inf (the CPU oracle restates it), under system zlib
(the reference's interop oracle, test/bin/simple.t) and under our own GPU inflate, consuming exactly `written` bytes."""
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import Error, Ok, _lib, batch, de, zl
from tests import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


def _inputs():
    rng = np.random.default_rng(1)
    txt = workloads.text(600000, 5)
    return [b"", b"a", b"abc", b"a" * 40, txt[:100], txt[:5000], txt[:65535], txt[:65536], txt[:65537], txt[:150000], txt,
            bytes(70000), bytes(300000), rng.integers(0, 256, 40000, dtype=np.uint8).tobytes(),
            rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(), (b"abcdefghij" * 7 + b"\n") * 3000,
            bytes(rng.integers(0, 4, 120000, dtype=np.uint8)), txt[:40000] + bytes(40000) + txt[40000:80000]]


def _roundtrip(inputs, res, fmt, eng):
    comps = []
    for k, (st, wr, comp) in enumerate(res):
        assert st == 0 and wr == len(comp) and wr <= _lib.lib().dcmp_deflate_bound(len(inputs[k]), fmt), k
        assert zlib.decompress(comp, 15 if fmt == batch.FMT_ZLIB else -15) == inputs[k], k
        rc, cons, prod, out = (oracle.zl_inflate_ns if fmt == batch.FMT_ZLIB else oracle.inflate_ns)(comp, len(inputs[k]) + 8)
        assert rc == 0 and cons == len(comp) and out == inputs[k], k
        comps.append(comp)
    back = batch.inflate_streams(comps, [len(x) for x in inputs], fmt=fmt, engine=eng)
    for k, (st, cons, prod, out) in enumerate(back):
        assert (st, cons, prod) == (0, len(comps[k]), len(inputs[k])) and out == inputs[k], k


@pytest.mark.parametrize("fmt", [batch.FMT_RAW, batch.FMT_ZLIB])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_roundtrip_vs_oracle_zlib_and_gpu_inflate(level, fmt, eng):
    inputs = _inputs()
    res = batch.deflate_streams(inputs, level=level, fmt=fmt, engine=eng)
    _roundtrip(inputs, res, fmt, eng)
    if level == 0:
        # level 0 = stored blocks (De.Def.Ns level 0: compress_none, de.ml:3935); test_ns.ml:1189-1200 pins the bytes
        assert batch.deflate_streams([b"\x01"], level=0, engine=eng)[0][2] == bytes.fromhex("010100feff01")


def test_compresses_text(eng):
    txt = workloads.text(1 << 20, 9)
    st, wr, comp = batch.deflate_streams([txt], level=6, fmt=batch.FMT_ZLIB, engine=eng)[0]
    ours, ref = len(txt) / wr, len(txt) / len(zlib.compress(txt, 6))
    assert ours > 0.8 * ref, (ours, ref)   # greedy parse, 64 KiB independent chunks: within 20 % of zlib -6


def test_ns_mirrors_and_errors(eng):
    txt = workloads.text(30000, 2)
    dst = de.bigstring_create(de.Def.Ns.compress_bound(len(txt)))
    r = de.Def.Ns.deflate(txt, dst, level=6, engine=eng)
    assert r.is_ok() and zlib.decompress(dst[:r.value].tobytes(), -15) == txt
    assert de.Def.Ns.deflate(txt, dst, level=13, engine=eng) == Error("Invalid_compression_level")   # de.ml:3930
    assert de.Def.Ns.deflate(txt, de.bigstring_create(100), level=6, engine=eng) == Error("Unexpected_end_of_output")
    zdst = de.bigstring_create(zl.Def.Ns.compress_bound(len(txt)))
    r = zl.Def.Ns.deflate(txt, zdst, level=6, engine=eng)
    assert r.is_ok() and zlib.decompress(zdst[:r.value].tobytes()) == txt
    assert zdst[:2].tobytes() == b"\x78\x9c"            # zl.ml:600-608, level 6 -> FLEVEL 2
    # De.Def.Ns -> De.Inf.Ns round trip with consumed == deflated length (test_ns.ml:1096-1187)
    back = de.bigstring_create(len(txt))
    assert de.Inf.Ns.inflate(dst[:de.Def.Ns.deflate(txt, dst, engine=eng).value], back, engine=eng).is_ok()
    assert back.tobytes() == txt


def test_many_small_inputs_in_one_batch(eng):
    rng = np.random.default_rng(4)
    inputs = [workloads.text(int(n), 100 + k) for k, n in enumerate(rng.integers(0, 9000, 500))]
    res = batch.deflate_streams(inputs, level=6, fmt=batch.FMT_ZLIB, engine=eng)
    _roundtrip(inputs, res, batch.FMT_ZLIB, eng)

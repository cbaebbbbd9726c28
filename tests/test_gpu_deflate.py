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


def test_window_edge_at_distance_32768_in_an_unaligned_batch(eng):
    """The far end of the window in streams that start at odd offsets of the batch: the compressor stages whole 16-byte
    granules by TMA and the ragged ends byte-wise; the bytes around a half are window, not padding
    (tests/cases.py: deflate_window_edge_stream)."""
    from tests import cases
    edge = cases.deflate_window_edge_stream()
    ins = [b"abc", edge, b"z" * 7, edge[5:], edge, bytes(9) + edge, workloads.text(200001, 9)]
    for level, align in ((1, 1), (6, 1), (9, 1), (6, 16)):   # (align = 1: the streams follow each other byte for byte)
        _roundtrip(ins, batch.deflate_streams(ins, level=level, fmt=batch.FMT_ZLIB, engine=eng, align=align), batch.FMT_ZLIB, eng)


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


def test_encoder_known_answers_on_the_gpu(eng):
    """SURVEY §8(f4) / VERDICT r1 item 7: De.T.make and De.Def.encode over a De.Queue, built ON THE DEVICE, against the
    reference's encoder known answers (test/test.ml:507-531 huffman_length_extra: 16 bytes; 1169-1237: tree_0 and the
    rfc5322 histogram — "the same tree as zlib") and against the CPU oracle's restatement on random input."""
    import json
    import os
    import random
    k = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "encoder.json")))["kats"]
    h = k["huffman_length_extra"]
    cmds = de.Queue.of_list([("Literal", 0), ("Literal", 0), ("Copy", 1, 258), ("Copy", 1, 256), "End"])
    assert cmds == h["cmds"]
    r = de.Def.encode_queue(cmds, last=h["last"], engine=eng)
    assert r.is_ok() and r.value.hex() == h["bytes"] == "edc101010000004020ff571b42c1ea04"
    t = k["tree_rfc5322"]
    lens, codes, mc = de.T.make(286, t["freqs"], engine=eng)
    assert [lens[ord("6")], codes[ord("6")]] == t["literal_6"] == [10, 0x35f] and [lens[256], codes[256]] == t["eob"] == [15, 0x6fff]
    assert (lens, codes, mc) == oracle.t_make(t["freqs"], 286)
    lens, codes, mc = de.T.make(286, [0] * 256 + [1], engine=eng)
    assert [[n, l] for n, l in enumerate(lens) if l] == k["tree_0"]["lengths"] and [[n, codes[n]] for n, l in enumerate(lens) if l] == k["tree_0"]["codes"]
    rng = random.Random(23)
    for length, maxl in ((286, 15), (30, 15), (19, 7)):
        for trial in range(20):
            f = [rng.choice([0, 0, 1, 2, 3, 50, 1000, rng.randrange(1 << 18)]) for _ in range(length)]
            if trial == 0:
                f = [1 << (j % 22) for j in range(length)]      # skewed: the overflow repair of gen_bitlen
            assert de.T.make(length, f, maxl, engine=eng) == oracle.t_make(f, length, maxl), (length, trial)
    data, cmds = bytearray(), []
    for _ in range(20000):
        if data and rng.random() < 0.45:
            off = rng.randint(1, min(len(data), 32768)); ln = rng.randint(3, 258)
            for j in range(ln):
                data.append(data[-off])
            cmds.append(de.Queue.copy(off, ln))
        else:
            b = rng.choice(b"abcdefgh \n") if rng.random() < 0.9 else rng.randrange(256)
            data.append(b); cmds.append(b)
    cmds.append(de.Queue.EOB)
    r = de.Def.encode_queue(cmds, engine=eng)
    want, _ = oracle.def_encode_dynamic(cmds, True)
    assert r.is_ok() and r.value == want and zlib.decompress(r.value, -15) == bytes(data)
    assert de.Def.encode_queue(cmds, cap=len(want) - 1, engine=eng) == Error("Unexpected_end_of_output")


def test_def_ns_known_answers_and_fixed_blocks(eng):
    """test/test_ns.ml:1189-1222: `encoder 0` (level 0, one stored block) and `encoder 1` (level 1: 56 zero bytes -> the five
    bytes 63 20 13 00 00: a Fixed block holding a literal, one match of 55 at distance 1, the end of block) byte for byte.
    Small inputs leave as Fixed blocks (Def.block_of_frequencies weighs Fixed against Dynamic, lib/de.ml:2411-2449)."""
    dst = de.bigstring_create(64)
    assert de.Def.Ns.deflate(b"\x01", dst, level=0, engine=eng) == Ok(6) and dst[:6].tobytes() == bytes.fromhex("010100feff01")
    assert de.Def.Ns.deflate(bytes(56), dst, level=1, engine=eng) == Ok(5) and dst[:5].tobytes() == bytes.fromhex("6320130000")
    small = [b"a", b"ab", b"abc", b"hello hello hello", workloads.text(15, 1), workloads.text(60, 2)]
    res = batch.deflate_streams(small, level=6, engine=eng)
    for x, (st, wr, comp) in zip(small, res):
        assert st == 0 and (comp[0] >> 1) & 3 == 1 and wr <= len(x) + 3, x       # BTYPE = 01
        assert zlib.decompress(comp, -15) == x and oracle.inflate_ns(comp, len(x))[:3] == (0, wr, len(x))

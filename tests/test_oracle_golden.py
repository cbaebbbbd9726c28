"""Pin the CPU oracle (oracle/*.c) to the reference's own golden vectors (CPU only).

The vectors were extracted from /root/reference/test/*.ml by tests/golden/make_golden.py.
"""
import json
import os
import zlib

import pytest

from tests import oracle

G = os.path.join(os.path.dirname(__file__), "golden")
BIG = 1 << 20  # test_ns.ml uses one module-level 0x100000-ish dst for most cases


def _load(fn):
    with open(os.path.join(G, fn)) as f:
        return json.load(f)["vectors"]


def _src(v):
    if "src_repeat" in v:
        a, n, tail = v["src_repeat"]
        return bytes.fromhex(a) * n + bytes.fromhex(tail)
    return bytes.fromhex(v["src"])


def _out(v):
    if "out_repeat" in v:
        a, n = v["out_repeat"]
        return bytes.fromhex(a) * n
    return bytes.fromhex(v["out"]) if "out" in v else None


@pytest.mark.parametrize("v", _load("inflate_ns.json"), ids=lambda v: v["name"])
def test_inflate_ns_golden(v):
    src = _src(v)
    cap = BIG if v["dst_len"] is None else v["dst_len"]
    rc, cons, prod, out = oracle.inflate_ns(src, cap)
    if "error" in v:
        assert oracle.NS_ERRORS[rc] == v["error"], v["cite"]
    else:
        assert rc == 0, (oracle.NS_ERRORS[rc], v["cite"])
        assert [cons, prod] == v["ok"], v["cite"]
        if _out(v) is not None:
            assert out == _out(v), v["cite"]


@pytest.mark.parametrize("v", _load("inflate_stream.json"), ids=lambda v: v["name"])
def test_inflate_stream_golden_via_ns(v):
    """The streaming KATs (test/test.ml) feed whole strings: output bytes / error class must
    agree with the one-shot path too (same algorithm, de.ml:1054-1261 vs 1667-1712)."""
    src = _src(v)
    rc, cons, prod, out = oracle.inflate_ns(src, BIG)
    if "malformed" in v:
        want = v["malformed"].replace(" ", "_")
        assert oracle.NS_ERRORS[rc].lower() == want.lower(), v["cite"]
    else:
        assert rc == 0, (oracle.NS_ERRORS[rc], v["cite"])
        assert out == _out(v), v["cite"]


@pytest.mark.parametrize("v", _load("inflate_ns.json") + _load("inflate_stream.json"), ids=lambda v: v["name"])
def test_inflate_matches_system_zlib(v):
    """fuzz/fuzz.ml:82-88: whatever system zlib accepts, De.Inf must decode identically."""
    src = _src(v)
    d = zlib.decompressobj(-15)
    try:
        want = d.decompress(src)
        ok = d.eof
    except zlib.error:
        ok = False
    rc, cons, prod, out = oracle.inflate_ns(src, BIG)
    if ok:
        assert rc == 0 and out == want
        assert cons == len(src) - len(d.unused_data)
    else:
        assert rc != 0 or "ok" in v or "out" in v  # zlib is stricter in a few documented places


@pytest.mark.parametrize("v", _load("lzo.json"), ids=lambda v: v["name"] + "@" + v["cite"].split("/")[-1])
def test_lzo_golden(v):
    src = _src(v)
    fn = oracle.lzo_uncompress_with_buffer if v["api"] == "uncompress_with_buffer" else oracle.lzo_uncompress
    rc, cons, prod, out = fn(src, 1 << 22)
    want = _out(v)
    if v["raw_hint"] == "error":
        assert rc != 0, v["cite"]
        if "error" in v:
            assert oracle.LZO_ERRORS[rc] == v["error"]
    else:
        assert rc == 0, (oracle.LZO_ERRORS[rc], v["cite"])
        if want is not None:
            assert out == want, v["cite"]


# ---- encoder side: De.T.make and Def's dynamic block (oracle/deflate_t.c) against the reference's known answers --------
def _enc():
    with open(os.path.join(os.path.dirname(__file__), "golden", "encoder.json")) as f:
        return json.load(f)["kats"]


def test_t_make_is_the_reference_tree():
    k = _enc()
    lens, codes, mc = oracle.t_make([0] * 256 + [1], 286)                       # tree_0 (test/test.ml:1169-1190)
    assert [[n, l] for n, l in enumerate(lens) if l] == k["tree_0"]["lengths"]
    assert [[n, codes[n]] for n, l in enumerate(lens) if l] == k["tree_0"]["codes"]
    r = k["tree_rfc5322"]                                                       # "the same tree as zlib" (test/test.ml:1192-1237)
    lens, codes, mc = oracle.t_make(r["freqs"], 286)
    assert [lens[ord("6")], codes[ord("6")]] == r["literal_6"] and [lens[256], codes[256]] == r["eob"]
    assert sum(2.0 ** -l for l in lens if l) == 1.0 and max(lens) == 15
    lens7, _, _ = oracle.t_make([5, 1, 1, 1, 1, 1, 1, 1, 40, 80, 160, 320, 640, 1280, 2560, 5120, 1, 1, 1], 19, 7)   # overflow repair to 7 bits
    assert max(lens7) == 7 and sum(2.0 ** -l for l in lens7 if l) == 1.0


def test_def_dynamic_block_bytes():
    k = _enc()["huffman_length_extra"]                                          # test/test.ml:507-531
    out, h = oracle.def_encode_dynamic(k["cmds"], k["last"])
    assert out.hex() == k["bytes"]
    assert zlib.decompress(out, -15) == bytes(2 + 258 + 256)
    # any command list: zlib and the oracle's own Inf take the block back
    import random
    rng = random.Random(5)
    data, cmds = bytearray(), []
    for _ in range(3000):
        if data and rng.random() < 0.4:
            off = rng.randint(1, min(len(data), 32768)); ln = rng.randint(3, 258)
            for j in range(ln):
                data.append(data[-off])
            cmds.append(0x2000000 | ((ln - 3) << 16) | (off - 1))
        else:
            b = rng.choice(b"abcdefgh \n") if rng.random() < 0.8 else rng.randrange(256)
            data.append(b); cmds.append(b)
    cmds.append(256)
    out, h = oracle.def_encode_dynamic(cmds, True)
    assert zlib.decompress(out, -15) == bytes(data)
    rc, cons, prod, got = oracle.inflate_ns(out, len(data))
    assert (rc, cons, prod) == (0, len(out), len(data)) and got == bytes(data)

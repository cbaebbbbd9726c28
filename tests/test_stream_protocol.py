"""The streaming state machines (decompress_b200/csrc/stream_core.h: the host side of dcmp_inf_* / dcmp_def_*) held to the
REFERENCE's own signal sequences and golden bytes, without a GPU: the same C++ templates are instantiated over the CPU
SIMT emulator, which runs the real device kernels (inflate_stream_kernel, the deflate kernels) on fibers.  Test
infrastructure only; tests/test_gpu_streaming.py runs the same cases through the product library on the B200.

Pinned here: test/test.ml:1450-1496 (git object, two `Manual chunks), 1536-1568 (`Flush, `Flush, `Flush, `End on "foo"),
1570-1657 (empty zlib: `End at once, 4-byte `o`, byte per byte), 1659-1758 (gzip: empty, FNAME, "foo", multiple flush),
1906-1921 (invalid FHCRC), 1960-1989 (FEXTRA, big-endian lengths), 97-767 (the streaming KATs)."""
import json
import os
import shutil
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import _lib, de, gz, stream, zl
from tests import oracle

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the emulator build needs g++")
G = os.path.join(os.path.dirname(__file__), "golden")
RAW, ZLIB, GZIP = _lib.FMT_RAW, _lib.FMT_ZLIB, _lib.FMT_GZIP


@pytest.fixture(scope="module")
def api():
    from tests import emu
    emu.build()
    return emu.StreamApi()


def _drain(d, o):
    """decode until `End / `Malformed: (signals, output)."""
    sig, out = [], bytearray()
    while True:
        r = d.decode()
        sig.append(r if isinstance(r, str) else r[0])
        if r == "Flush":
            out += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        elif r == "End":
            out += o[:len(o) - d.dst_rem()].tobytes()
            return sig, bytes(out), None
        elif r == "Await":
            raise AssertionError("unexpected `Await")
        else:
            return sig, bytes(out), r[1]


def _vectors():
    with open(os.path.join(G, "inflate_stream.json")) as f:
        return json.load(f)["vectors"]


@pytest.mark.parametrize("v", _vectors(), ids=lambda v: v["name"])
def test_stream_kats(v, api):
    o = de.bigstring_create(de.io_buffer_size)
    d = de.Inf.decoder(("String", bytes.fromhex(v["src"])), o, de.make_window(bits=15), api=api)
    sig, out, err = _drain(d, o)
    if "malformed" in v:
        assert err == v["malformed"], v["cite"]
        assert d.decode() == "End"          # the decoder is stuck afterwards (continuation `final`, de.ml:700)
    else:
        assert err is None, v["cite"]
    if "out" in v and "malformed" not in v:
        assert out == bytes.fromhex(v["out"]), v["cite"]


def test_multiple_flush_zlib(api):
    """test/test.ml:1536-1568: `Flush until flush is called, the contents with the third, then `End."""
    comp = zlib.compress(b"foo", 0)
    o = de.bigstring_create(de.io_buffer_size)
    d = zl.Inf.decoder(("String", comp), o, api=api)
    assert d.decode() == "Flush"
    assert d.decode() == "Flush"
    assert d.decode() == "Flush"
    assert o[:len(o) - d.dst_rem()].tobytes() == b"foo"
    d.flush()
    assert d.decode() == "End"
    assert d.write() == 3 and d.checksum() == zlib.adler32(b"foo")


def test_empty_zlib_end_at_once_small_output_and_byte_per_byte(api):
    """test/test.ml:1570-1657."""
    empty = zlib.compress(b"", 3)
    for size in (de.io_buffer_size, 4):
        o = de.bigstring_create(size)
        d = zl.Inf.decoder(("String", empty), o, api=api)
        assert d.decode() == "End" and len(o) - d.dst_rem() == 0
    o = de.bigstring_create(de.io_buffer_size)
    d = zl.Inf.decoder("Manual", o, api=api)
    i = np.zeros(1, dtype=np.uint8)
    k, awaits = 0, 0
    while True:
        r = d.decode()
        assert r != "Flush"
        if r == "Await":
            awaits += 1
            if k < len(empty):
                i[0] = empty[k]
                d.src(i, 0, 1)
                k += 1
            else:
                d.src(i, 0, 0)
        else:
            break
    assert r == "End" and len(o) - d.dst_rem() == 0 and awaits == len(empty)


GIT_OBJECT = [b"\x78\x01\x01\x1e\x00\xe1\xff", b"\x9d\x02\x9d\x02\x90\xad\x14", b"\x72\xb9\xb4\x44\x59\x5d\x21",
              b"\x05\xb7\x34\x4f\x64\xe9\xa8", b"\x5a\x3e\xd9\x91\x1f\x44\x91", b"\xc1\x5c\xbd\xe5\x0c\xd1"]


def test_git_object_two_manual_chunks(api):
    """test/test.ml:1450-1496: a stored block split over two `Manual inputs: `Await exactly once, after the first."""
    o = de.bigstring_create(de.io_buffer_size)
    whole = b"".join(GIT_OBJECT)
    sig, out, err = _drain(zl.Inf.decoder(("String", whole), o, api=api), o)
    assert err is None and out == zlib.decompress(whole)
    d = zl.Inf.decoder("Manual", o, api=api)
    i0, i1 = GIT_OBJECT[0], b"".join(GIT_OBJECT[1:])
    d.src(i0, 0, len(i0))
    seen, got = [], bytearray()
    while True:
        r = d.decode()
        seen.append(r)
        if r == "Await":
            assert seen.count("Await") == 1, "`Await after the second chunk"
            d.src(i1, 0, len(i1))
        elif r == "Flush":
            got += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        else:
            break
    assert r == "End" and bytes(got) == zlib.decompress(whole)


def test_output_before_the_end_of_input_and_bounded_memory(api):
    """What round 1's facade could not do: output appears while input is still arriving, and what the decoder holds is
    bounded whatever the stream's length (input tail + one launch's output, <= ~1 MiB)."""
    txt = workloads.text(3 << 20, 21)
    comp = zlib.compress(txt, 6)
    o = de.bigstring_create(de.io_buffer_size)
    d = zl.Inf.decoder("Manual", o, api=api)
    pos, out, first_flush_at, peak = 0, bytearray(), None, 0
    while True:
        r = d.decode()
        peak = max(peak, api.inf_buffered(d._h))
        if r == "Await":
            n = min(65536, len(comp) - pos)
            d.src(comp, pos, n)
            pos += n
        elif r == "Flush":
            if first_flush_at is None:
                first_flush_at = pos
            out += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        else:
            break
    assert r == "End" and bytes(out) == txt
    assert first_flush_at is not None and first_flush_at <= 2 * 65536 < len(comp)
    assert peak <= (1 << 20) + (1 << 17)
    assert d.launches() <= len(comp) // 65536 + 6      # one launch per input chunk (+ the output-budget stops)


def test_de_inf_truncated_and_trailing_input(api):
    txt = workloads.text(70000, 2)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(txt) + co.flush()
    o = de.bigstring_create(1 << 17)
    sig, out, err = _drain(de.Inf.decoder(("String", comp[:-7]), o, api=api), o)
    assert err == "Unexpected end of input" and txt.startswith(out)
    d = de.Inf.decoder(("String", comp + b"trailer!"), o, api=api)
    sig, out, err = _drain(d, o)
    assert err is None and out == txt and d.src_rem() == 8   # what follows the stream is given back (de.ml:1459-1466)
    assert d.checksum() == zlib.adler32(txt)
    d.reset()                                                # De.Inf.reset: the same decoder decodes again
    sig, out, err = _drain(d, o)
    assert err is None and out == txt


GZ_EMPTY = bytes.fromhex("1f8b080082e5535e0003030000000000" "00000000")
GZ_EMPTY_NAME = bytes.fromhex("1f8b0808a6ec535e0003746573740003" "000000000000000000")
GZ_FOO = bytes.fromhex("1f8b08082df1535e0003666f6f004bcb" "cf07002165738c03000000")
GZ_BAD_HCRC = bytes.fromhex("1f8b080a000000000203666f6f2e677a" "00" "4bcb71de" "cf575053484a2c02008f47" "e4dd09000000")
GZ_EXTRA = bytes.fromhex("1f8b080c63cb5f5e0003" "000a" "6c7800067562756e7475" "666f6f2e7478" "74004bcbcfe70200a865327e04000000")


def test_gzip_reference_vectors(api):
    """test/test.ml:1659-1758, 1906-1921, 1960-1989."""
    o = de.bigstring_create(de.io_buffer_size)
    d = gz.Inf.decoder(("String", GZ_EMPTY), o, api=api)
    assert d.decode() == "End" and d.dst_rem() == len(o) and d.os() == 3
    d = gz.Inf.decoder(("String", GZ_EMPTY_NAME), o, api=api)
    assert d.decode() == "End" and d.dst_rem() == len(o) and d.filename() == b"test"
    d = gz.Inf.decoder(("String", GZ_FOO), o, api=api)
    assert d.decode() == "Flush" and d.decode() == "Flush" and d.decode() == "Flush"     # multiple flush GZip
    assert o[:len(o) - d.dst_rem()].tobytes() == b"foo" and d.filename() == b"foo"
    d.flush()
    assert d.decode() == "End" and d.write() == 3 and d.checksum() == zlib.crc32(b"foo")
    d = gz.Inf.decoder(("String", GZ_BAD_HCRC), o, api=api)
    assert d.decode() == ("Malformed", "Invalid GZip header checksum")
    d = gz.Inf.decoder(("String", GZ_EXTRA), o, api=api)
    assert d.decode() == "Flush" and o[:len(o) - d.dst_rem()].tobytes() == b"foo\n"
    d.flush()
    assert d.decode() == "End" and d.extra("lx") == b"ubuntu" and d.extra("zz") is None and d.filename() == b"foo.txt"
    bad = bytearray(GZ_FOO); bad[-6] ^= 1
    sig, out, err = _drain(gz.Inf.decoder(("String", bytes(bad)), o, api=api), o)
    assert err.startswith("Invalid checksum (expect:")
    bad = bytearray(GZ_FOO); bad[-4] = 4
    sig, out, err = _drain(gz.Inf.decoder(("String", bytes(bad)), o, api=api), o)
    assert err == "Invalid input size (expect:4, inflated:3)"
    sig, out, err = _drain(gz.Inf.decoder(("String", b"\x1f\x8c" + GZ_FOO[2:]), o, api=api), o)
    assert err == "Invalid GZip header"
    sig, out, err = _drain(gz.Inf.decoder(("String", GZ_FOO[:-3]), o, api=api), o)
    assert err == "Unexpected end of input"
    # the header arrives byte by byte (the reference needs its ten fixed bytes in one piece, gz.ml:489-492; here any split works)
    d = gz.Inf.decoder("Manual", o, api=api)
    k, out = 0, bytearray()
    while True:
        r = d.decode()
        if r == "Await":
            d.src(GZ_EXTRA, k, 1 if k < len(GZ_EXTRA) else 0)
            k += 1
        elif r == "Flush":
            out += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        else:
            break
    assert r == "End" and bytes(out) == b"foo\n" and d.extra("lx") == b"ubuntu"


def test_multi_member_gzip_sequential_loop(api):
    """lib/gz.ml:550, 582-610: `End after one member, src_rem says where the next one starts, reset, go on."""
    members = workloads.c4_members(600000, seed=5, lo=20000, hi=120000)
    blob = b"".join(m for m, _ in members)
    res, outs = gz.inflate_multi_sequential(blob, api=api, o_size=65536)
    assert res.is_ok() and res.value == len(members)
    assert outs == [p for _, p in members]


def _encode(e, data, o_size, chunk):
    """The `Manual / `Manual loop of test/test.ml:1597-1625: (signals, bytes)."""
    o = np.zeros(o_size, dtype=np.uint8)
    e.dst(o, 0, o_size)
    pos, out, sig = 0, bytearray(), []
    while True:
        r = e.encode()
        sig.append(r)
        if r == "Await":
            n = min(chunk, len(data) - pos)
            e.src(data, pos, n)
            pos += n
        elif r == "Flush":
            out += o[:o_size - e.dst_rem()].tobytes()
            e.dst(o, 0, o_size)
        else:
            assert o_size - e.dst_rem() == 0, "`End with bytes left in `o` (test.ml:1618-1622)"
            return sig, bytes(out)


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_zl_def_encoder_streams(level, api):
    txt = workloads.text(700000, 13)       # several 256 KiB segments (emu.StreamApi(seg_bytes))
    for data, o_size, chunk in ((txt, 65536, 65536), (txt[:100], 4, 7), (b"", 4, 1), (b"a", 1, 1)):
        sig, comp = _encode(zl.Def.encoder(None, None, level, api=api), data, o_size, chunk)
        assert zlib.decompress(comp) == data
        rc, cons, prod, out = oracle.zl_inflate_ns(comp, len(data) + 8)
        assert (rc, cons, prod) == (0, len(comp), len(data)) and out == data     # the reference's Inf takes it back
        assert sig[-1] == "End" and "Await" in sig
    with pytest.raises(ValueError):
        zl.Def.encoder(None, None, 10, api=api)


def test_gz_def_encoder_header_bytes_and_roundtrip(api):
    import gzip
    txt = workloads.text(300000, 14)
    e = gz.Def.encoder(mtime=0x5e53f12d, os="Unix", level=6, filename=b"foo", api=api)
    sig, comp = _encode(e, txt, 65536, 65536)
    assert comp[:14] == bytes.fromhex("1f8b0808" "5e53f12d" "0003" "666f6f00")     # MTIME big-endian, as lib/gz.ml:797
    assert gzip.decompress(comp) == txt
    assert comp[-8:] == zlib.crc32(txt).to_bytes(4, "little") + (len(txt) & 0xFFFFFFFF).to_bytes(4, "little")
    # FHCRC: what Gz.Def writes, Gz.Inf accepts (test.ml:1923-1958, every OS)
    for os_name in ("FAT", "Unix", "NTFS"):
        e = gz.Def.encoder(mtime=0, os=os_name, level=4, hcrc=True, comment=b"c", api=api)
        sig, comp = _encode(e, txt[:256], 65536, 65536)
        o = de.bigstring_create(65536)
        d = gz.Inf.decoder(("String", comp), o, api=api)
        s2, out, err = _drain(d, o)
        assert err is None and out == txt[:256] and gz.os_of_int(d.os()) == os_name and d.comment() == b"c"
    sig, comp = _encode(gz.Def.encoder(mtime=0, os="Unix", level=3, api=api), b"", 65536, 1)   # test.ml:1760-1785
    o = de.bigstring_create(65536)
    d = gz.Inf.decoder(("String", comp), o, api=api)
    assert d.decode() == "End" and d.dst_rem() == len(o)


def test_higher_loops(api):
    txt = workloads.text(400000, 15)
    i, o = de.bigstring_create(de.io_buffer_size), de.bigstring_create(de.io_buffer_size)

    def reader(data):
        pos = [0]

        def refill(buf):
            n = min(len(buf), len(data) - pos[0])
            buf[:n] = np.frombuffer(data, dtype=np.uint8)[pos[0]:pos[0] + n]
            pos[0] += n
            return n
        return refill

    for mod, fmt_wbits in ((zl, 15), (de, -15)):
        comp = bytearray()
        mod.Higher.compress(None, None, reader(txt), lambda b, n: comp.extend(b[:n].tobytes()), i, o, level=6, api=api)
        assert zlib.decompress(bytes(comp), fmt_wbits) == txt
        back = bytearray()
        first = (lambda bits: None) if mod is zl else None
        r = mod.Higher.uncompress(first, reader(bytes(comp)), lambda b, n: back.extend(b[:n].tobytes()), i, o, api=api)
        assert r.is_ok() and bytes(back) == txt
    comp = bytearray()
    gz.Higher.compress(None, None, reader(txt), lambda b, n: comp.extend(b[:n].tobytes()), 0, {"os": 3}, i, o, level=6, filename=b"x", api=api)
    back = bytearray()
    r = gz.Higher.uncompress(reader(bytes(comp)), lambda b, n: back.extend(b[:n].tobytes()), i, o, api=api)
    assert r.is_ok() and r.value == {"filename": b"x", "comment": None, "os": "Unix"} and bytes(back) == txt
    bad = bytearray(zlib.compress(txt)); bad[-1] ^= 1
    r = zl.Higher.uncompress(None, reader(bytes(bad)), lambda b, n: None, i, o, api=api)
    assert not r.is_ok() and r.error[1].startswith("Invalid checksum (expect:")

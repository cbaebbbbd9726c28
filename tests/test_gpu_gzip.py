"""GPU parity tests for gzip members (host framing per lib/gz.ml + DEFLATE body on the GPU): the reference's gzip
goldens (test/test.ml:1659-1758, 1906-1921, 1960-1989), Python's gzip module as the interop oracle, error classes."""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import Error, Ok, _lib, batch, gz
from tests import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


def _member(plain, fname=None, comment=None, extra=None, hcrc=False, mtime=0):
    """RFC1952 member the way lib/gz.ml writes/reads it: MTIME big-endian (gz.ml:479, 801), XLEN big-endian
    (gz.ml:455), FHCRC = upper 16 bits of CRC-32 over fixed header + FNAME + FCOMMENT, big-endian (gz.ml:422-440)."""
    flg = (4 if extra is not None else 0) | (8 if fname is not None else 0) | (16 if comment is not None else 0) | (2 if hcrc else 0)
    fixed = b"\x1f\x8b\x08" + bytes([flg]) + struct.pack(">I", mtime) + b"\x00\x03"
    out = fixed
    crc_src = fixed
    if extra is not None:
        out += struct.pack(">H", len(extra)) + extra
    for f in (fname, comment):
        if f is not None:
            out += f + b"\x00"
            crc_src += f + b"\x00"
    if hcrc:
        out += struct.pack(">H", (zlib.crc32(crc_src) >> 16) & 0xffff)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(plain) + co.flush()
    return out + body + struct.pack("<II", zlib.crc32(plain) & 0xffffffff, len(plain) & 0xffffffff)


def test_reference_gzip_goldens(eng):
    # test/test.ml:1703-1704: "foo" -> CRC 0x8c736521, ISIZE 3
    foo = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + bytes.fromhex("4bcbcf0700") + bytes.fromhex("2165738c") + struct.pack("<I", 3)
    dst = np.zeros(64, dtype=np.uint8)
    assert gz.inflate_member(foo, dst, engine=eng) == Ok((len(foo), 3)) and dst[:3].tobytes() == b"foo"
    # FNAME (test.ml:1727-1758), FEXTRA with big-endian XLEN (test.ml:1960-1989), FHCRC ok / invalid (test.ml:1906-1921)
    txt = workloads.text(3000, 1)
    for kw in (dict(fname=b"foo.txt"), dict(extra=b"lx\x00\x06abcdef"), dict(fname=b"a", comment=b"hello", hcrc=True), dict(hcrc=True)):
        m = _member(txt, **kw)
        dst = np.zeros(4000, dtype=np.uint8)
        assert gz.inflate_member(m, dst, engine=eng) == Ok((len(m), len(txt))), kw
        assert dst[:len(txt)].tobytes() == txt
    bad = bytearray(_member(txt, fname=b"a", hcrc=True))
    bad[13] ^= 0xff  # FHCRC bytes
    assert gz.inflate_member(bytes(bad), np.zeros(4000, dtype=np.uint8), engine=eng) == Error("Invalid GZip header checksum")


def test_error_classes(eng):
    txt = workloads.text(5000, 3)
    m = _member(txt)
    dst = lambda: np.zeros(6000, dtype=np.uint8)
    assert gz.inflate_member(b"\x1f\x8c" + m[2:], dst(), engine=eng) == Error("Invalid GZip header")
    assert gz.inflate_member(m[:7], dst(), engine=eng) == Error("Unexpected end of input")
    assert gz.inflate_member(m[:-3], dst(), engine=eng) == Error("Unexpected end of input")
    c = bytearray(m); c[-8] ^= 1
    assert gz.inflate_member(bytes(c), dst(), engine=eng) == Error("Invalid checksum")
    c = bytearray(m); c[-1] ^= 1
    assert gz.inflate_member(bytes(c), dst(), engine=eng) == Error("Invalid input size")
    c = bytearray(m); c[40] ^= 0x55
    assert not gz.inflate_member(bytes(c), dst(), engine=eng).is_ok()


def test_python_gzip_interop_and_multi_member_c4(eng):
    """BASELINE config C4 at reduced size: concatenated members (16..256 KiB payloads), sharded by a sidecar index."""
    members = workloads.c4_members(3 << 20, seed=2)
    blob = b"".join(m for m, _ in members)
    assert gzip.GzipFile(fileobj=io.BytesIO(blob)).read() == b"".join(p for _, p in members)
    off, pos = [], 0
    for m, _ in members:
        off.append(pos)
        pos += len(m)
    res = gz.inflate_multi(np.frombuffer(blob, dtype=np.uint8), off, [len(p) for _, p in members], engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod) == (0, len(members[k][0]), len(members[k][1])) and out == members[k][1], k
    # without an index: the reference's End -> src_rem -> reset loop
    r, outs = gz.inflate_multi_sequential(np.frombuffer(blob[:off[4]], dtype=np.uint8), 256 << 10, engine=eng)
    assert r == Ok(4) and outs == [p for _, p in members[:4]]
    # a member written by Python's gzip module (MTIME little-endian there; the field is not interpreted on decode)
    py = gzip.compress(members[0][1], 6)
    d = np.zeros(len(members[0][1]), dtype=np.uint8)
    assert gz.inflate_member(py, d, engine=eng) == Ok((len(py), len(members[0][1]))) and d.tobytes() == members[0][1]


def test_gzip_deflate_batch_members_on_the_device(eng):
    """DCMP_FMT_GZIP in dcmp_deflate_batch (Gz.Def's framing, lib/gz.ml:794-813, 725-733, written on the device): Python's
    gzip module, the oracle's CRC-32 and our own gzip inflate take the members back."""
    txt = workloads.text(900000, 31)
    ins = [txt, b"", b"a", txt[:65536], txt[:70001], bytes(200000)]
    for level in (0, 6, 9):
        res = batch.deflate_streams(ins, level=level, fmt=_lib.FMT_GZIP, engine=eng)
        comps = []
        for k, (st, wr, comp) in enumerate(res):
            assert st == 0 and wr == len(comp) <= _lib.lib().dcmp_deflate_bound(len(ins[k]), _lib.FMT_GZIP), (level, k)
            assert comp[:10] == bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 2 if level >= 9 else 0, 3])
            assert gzip.decompress(comp) == ins[k], (level, k)
            assert comp[-8:] == (oracle.crc32(ins[k]) & 0xFFFFFFFF).to_bytes(4, "little") + len(ins[k]).to_bytes(4, "little")
            comps.append(comp)
        back = batch.inflate_streams(comps, [len(x) for x in ins], fmt=_lib.FMT_GZIP, engine=eng)
        for k, (st, cons, prod, out) in enumerate(back):
            assert (st, cons, prod) == (0, len(comps[k]), len(ins[k])) and out == ins[k], (level, k)
    st, wr, comp = batch.deflate_streams([txt], [17], level=6, fmt=_lib.FMT_GZIP, engine=eng)[0]
    assert st == 2 and wr == 0      # Unexpected_end_of_output: no room for the framing


def _bgzf(data, block=60000):
    """A BGZF file as bgzip writes it: members of <= 64 KiB with the 'BC' size sub-field (XLEN little-endian), an empty EOF member."""
    out = bytearray()
    for pos in list(range(0, len(data), block)) + [None]:
        chunk = b"" if pos is None else data[pos:pos + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out += bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff]) + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out)


def test_bgzf_member_index_and_batch_inflate(eng):
    """a multi-member file cut into units WITHOUT decoding — BGZF's own size fields (dcmp_gz_bgzf_index) — and
    inflated as one batch; XLEN is little-endian there (DCMP_GZIP_RFC_XLEN), where the reference reads it big-endian."""
    txt = workloads.text(1_000_000, 77)
    blob = _bgzf(txt)
    assert gzip.decompress(blob) == txt                   # what every other tool makes of it
    off, ln, isz = gz.bgzf_index(blob)
    assert len(off) == (len(txt) + 59999) // 60000 + 1 and int(isz.sum()) == len(txt) and off[0] == 0 and off[-1] + ln[-1] == len(blob)
    assert gz.inflate_bgzf(blob, engine=eng) == Ok(txt)
    bad = bytearray(blob); bad[int(off[3]) + 40] ^= 0x20
    assert not gz.inflate_bgzf(bytes(bad), engine=eng).is_ok()
    assert gz.bgzf_index(gzip.compress(txt[:1000])) is None   # a plain gzip member carries no size
    # the reference's reading of the same bytes (XLEN big-endian, lib/gz.ml:455): 0x0600 bytes of "extra", then garbage
    st, cons, prod, out = gz.inflate_members([blob[:int(ln[0])]], [int(isz[0])], engine=eng)[0]
    assert st != 0

"""The streaming facade (De.Inf.decoder / De.Higher over the batch engine) against the reference's streaming KATs
(test/test.ml:97-300, 302-767: input -> output bytes or `Malformed message) and its signal protocol."""
import json
import os
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import Ok, _lib, de, stream
from tests import test_stream_protocol as P

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


@pytest.fixture(scope="module")
def api(eng):
    """The product's entry points (dcmp_inf_* / dcmp_def_* on the B200) for the protocol cases of test_stream_protocol.py."""
    return stream.ProductApi(eng)


# the reference's sequences and golden bytes (test/test.ml:1450-1496, 1536-1758, 1906-1989), through libdcmp_b200.so
test_kats_c_abi = P.test_stream_kats
test_multiple_flush_zlib = P.test_multiple_flush_zlib
test_empty_zlib = P.test_empty_zlib_end_at_once_small_output_and_byte_per_byte
test_git_object_two_manual_chunks = P.test_git_object_two_manual_chunks
test_output_before_the_end_of_input_and_bounded_memory = P.test_output_before_the_end_of_input_and_bounded_memory
test_de_inf_truncated_and_trailing_input = P.test_de_inf_truncated_and_trailing_input
test_gzip_reference_vectors = P.test_gzip_reference_vectors
test_multi_member_gzip_sequential_loop = P.test_multi_member_gzip_sequential_loop
test_zl_def_encoder_streams = P.test_zl_def_encoder_streams
test_gz_def_encoder_header_bytes_and_roundtrip = P.test_gz_def_encoder_header_bytes_and_roundtrip
test_higher_loops = P.test_higher_loops


def _vectors():
    with open(os.path.join(G, "inflate_stream.json")) as f:
        return json.load(f)["vectors"]


@pytest.mark.parametrize("v", _vectors(), ids=lambda v: v["name"])
def test_stream_goldens_through_the_decoder(v, eng):
    o = de.bigstring_create(de.io_buffer_size)
    d = de.Inf.decoder(("String", bytes.fromhex(v["src"])), o, de.make_window(bits=15), engine=eng)
    out = bytearray()
    while True:
        r = d.decode()
        if r == "Flush":
            out += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        elif r == "End":
            out += o[:len(o) - d.dst_rem()].tobytes()
            assert "malformed" not in v, v["cite"]
            break
        else:
            assert r[0] == "Malformed" and r[1] == v["malformed"], (r, v["cite"])
            break
    if "out" in v and "malformed" not in v:
        assert bytes(out) == bytes.fromhex(v["out"]), v["cite"]


def test_signal_protocol_manual_source_small_output_buffer(eng):
    """test/test.ml:1515-1568: byte-per-byte feeding, a 4-byte output buffer, idempotent `Flush."""
    txt = workloads.text(1000, 4)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(txt) + co.flush()
    o = de.bigstring_create(4)
    d = de.Inf.decoder("Manual", o, de.make_window(bits=15), engine=eng)
    out, sigs = bytearray(), []
    k = 0
    while True:
        r = d.decode()
        sigs.append(r if isinstance(r, str) else r[0])
        if r == "Await":
            d.src(comp, k, 1 if k < len(comp) else 0)   # one byte at a time, then end of input
            k += 1
        elif r == "Flush":
            assert d.decode() == "Flush"                # nothing moves until flush (zl.ml:288-291)
            out += o[:4 - d.dst_rem()].tobytes()
            d.flush()
        else:
            out += o[:4 - d.dst_rem()].tobytes()
            break
    assert r == "End" and bytes(out) == txt
    # one `Await per input byte — none for the end of the input: the final block ends with the last byte, as in the reference
    assert sigs.count("Await") == len(comp) and sigs.count("Flush") == 250 - 1 and sigs[-1] == "End"
    assert d.checksum() == zlib.adler32(txt)


def test_higher_uncompress_and_of_string(eng):
    txt = workloads.text(300000, 8)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(txt) + co.flush()
    i, o = de.bigstring_create(de.io_buffer_size), de.bigstring_create(de.io_buffer_size)
    pos, out = [0], bytearray()

    def refill(buf):
        n = min(len(buf), len(comp) - pos[0])
        buf[:n] = np.frombuffer(comp, dtype=np.uint8)[pos[0]:pos[0] + n]
        pos[0] += n
        return n

    def flush(buf, n):
        out.extend(buf[:n].tobytes())

    assert de.Higher.uncompress(de.make_window(bits=15), refill, flush, i, o, engine=eng) == Ok(None)
    assert bytes(out) == txt
    out.clear()
    assert de.Higher.of_string(o, de.make_window(bits=15), comp, flush, engine=eng) == Ok(None) and bytes(out) == txt
    r = de.Higher.of_string(o, de.make_window(bits=15), b"\x06", flush, engine=eng)
    assert not r.is_ok() and r.error == ("Msg", "Invalid kind of block")


def test_higher_compress_roundtrips(eng):
    """Zl.Higher.compress -> Zl.Higher.uncompress and De.Higher.compress / to_string -> system zlib (bench/b.ml's loop)."""
    from decompress_b200 import zl
    txt = workloads.text(500000, 12)
    i, o = de.bigstring_create(de.io_buffer_size), de.bigstring_create(de.io_buffer_size)

    def reader(data):
        pos = [0]

        def refill(buf):
            n = min(len(buf), len(data) - pos[0])
            buf[:n] = np.frombuffer(data, dtype=np.uint8)[pos[0]:pos[0] + n]
            pos[0] += n
            return n
        return refill

    comp = bytearray()
    zl.Higher.compress(None, None, reader(txt), lambda buf, n: comp.extend(buf[:n].tobytes()), i, o, level=6, engine=eng)
    assert zlib.decompress(bytes(comp)) == txt
    back = bytearray()
    r = zl.Higher.uncompress(lambda bits: de.make_window(bits), reader(bytes(comp)), lambda buf, n: back.extend(buf[:n].tobytes()), i, o, engine=eng)
    assert r == Ok(None) and bytes(back) == txt
    raw = de.Higher.to_string(None, None, reader(txt), i, engine=eng)
    assert zlib.decompress(raw, -15) == txt
    bad = bytearray(comp); bad[-1] ^= 1
    r = zl.Higher.uncompress(None, reader(bytes(bad)), lambda buf, n: None, i, o, engine=eng)
    assert not r.is_ok() and r.error[1].startswith("Invalid checksum")

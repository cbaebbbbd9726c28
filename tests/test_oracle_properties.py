"""CPU: the oracle against system zlib on generated streams (fuzz/fuzz.ml:82-88 property) and on
the reference's corpus when it is present (test/test.ml:1239-1448 round-trips, decoder side)."""
import os
import zlib

import pytest

from tests import cases, oracle

CORPUS = "/root/reference/test/corpus"


@pytest.mark.parametrize("name,comp,plain", cases.valid_streams(), ids=lambda v: v if isinstance(v, str) else "")
def test_oracle_inflates_zlib_output(name, comp, plain):
    rc, cons, prod, out = oracle.inflate_ns(comp, len(plain) + 16)
    assert rc == 0 and out == plain and cons == len(comp)
    # exact-size destination works, one byte less fails with Unexpected_end_of_output
    rc, _, _, out = oracle.inflate_ns(comp, len(plain))
    assert rc == 0 and out == plain
    if plain:
        rc, _, _, _ = oracle.inflate_ns(comp, len(plain) - 1)
        assert oracle.NS_ERRORS[rc] == "Unexpected_end_of_output"


def test_oracle_vs_zlib_on_mutations():
    agree = 0
    for name, s in cases.mutated_streams(400):
        d = zlib.decompressobj(-15)
        try:
            want = d.decompress(s)
            ok = d.eof
        except zlib.error:
            ok = False
        rc, cons, prod, out = oracle.inflate_ns(s, 1 << 20)
        if ok:  # whatever zlib accepts the reference must decode identically
            assert rc == 0 and out == want and cons == len(s) - len(d.unused_data), name
            agree += 1
    assert agree > 20


@pytest.mark.skipif(not os.path.isdir(CORPUS), reason="reference corpus not present on this box")
@pytest.mark.parametrize("level", [1, 6, 9])
def test_oracle_on_reference_corpus(level):
    for fn in sorted(os.listdir(CORPUS)):
        data = open(os.path.join(CORPUS, fn), "rb").read()
        comp = cases.deflate_raw(data, level)
        rc, cons, prod, out = oracle.inflate_ns(comp, len(data))
        assert rc == 0 and out == data and cons == len(comp), fn
        z = zlib.compress(data, level)
        rc, cons, prod, out = oracle.zl_inflate_ns(z, len(data))
        assert rc == 0 and out == data and cons == len(z), fn


def test_zl_ns_errors():
    z = bytearray(zlib.compress(b"hello hello hello hello"))
    rc, *_ = oracle.zl_inflate_ns(bytes(z), 100)
    assert rc == 0
    bad = bytearray(z); bad[-1] ^= 1
    assert oracle.NS_ERRORS[oracle.zl_inflate_ns(bytes(bad), 100)[0]] == "Invalid_checksum"
    bad = bytearray(z); bad[0] = 0x79
    assert oracle.NS_ERRORS[oracle.zl_inflate_ns(bytes(bad), 100)[0]] == "Invalid_header"
    assert oracle.NS_ERRORS[oracle.zl_inflate_ns(b"\x78", 100)[0]] == "Unexpected_end_of_input"

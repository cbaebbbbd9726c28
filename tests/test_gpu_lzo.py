"""GPU parity tests for the LZO1X decode path (dcmp_lzo_uncompress_batch through the C ABI) against the
CPU oracle, the reference's golden vectors (test/test_lzo.ml, test/test.ml) and the reference's own
differential oracle, vendored minilzo.  Bit-exact: status, produced and every output byte."""
import json
import os
import random

import numpy as np
import pytest

import workloads
from decompress_b200 import Error, Ok, _lib, lzo
from tests import cases, oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
_ST = {0: 0, 1: 16, 2: 17, 3: 18}  # oracle code -> include/dcmp.h status


def _vectors():
    with open(os.path.join(G, "lzo.json")) as f:
        return json.load(f)["vectors"]


def _src(v):
    if "src" in v:
        return bytes.fromhex(v["src"])
    pre, times, post = v["src_repeat"]
    return bytes.fromhex(pre) * times + bytes.fromhex(post)


def _out(v):
    if "out" in v:
        return bytes.fromhex(v["out"])
    if "out_repeat" in v:
        pre, times = v["out_repeat"]
        return bytes.fromhex(pre) * times
    return None


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


@pytest.mark.parametrize("v", _vectors(), ids=lambda v: v["name"] + "@" + v["cite"].split("/")[-1])
def test_lzo_golden(v, eng):
    """The reference's KATs, called the way test/test_lzo.ml calls them."""
    src = _src(v)
    if v["api"] == "uncompress_with_buffer":
        res = lzo.uncompress_with_buffer(src, engine=eng)
        got = res.value if res.is_ok() else None
    else:
        out = np.zeros(1 << 22, dtype=np.uint8)
        res = lzo.uncompress(src, out, engine=eng)
        got = res.value.tobytes() if res.is_ok() else None
    if v["raw_hint"] == "error":
        assert not res.is_ok(), v["cite"]
        if "error" in v:
            assert res == Error(v["error"]), v["cite"]
    else:
        assert res.is_ok(), (res, v["cite"])
        want = _out(v)
        if want is not None:
            assert got == want, v["cite"]


def _mutants(comps, count, seed):
    rnd = random.Random(seed)
    out = []
    for _ in range(count):
        s = bytearray(rnd.choice(comps))
        m = rnd.randrange(4)
        if m == 0:
            s = s[:rnd.randrange(1, len(s))]
        elif m == 1:
            s[rnd.randrange(len(s))] ^= 1 << rnd.randrange(8)
        elif m == 2:
            s = bytearray(rnd.randrange(256) for _ in range(rnd.randrange(1, 100)))
        else:
            s[rnd.randrange(min(40, len(s)))] ^= 1 << rnd.randrange(8)
        out.append(bytes(s))
    return out


def _check(blocks, caps, mode, eng):
    fn = oracle.lzo_uncompress if mode == 0 else oracle.lzo_uncompress_with_buffer
    res = lzo.uncompress_batch(blocks, caps, mode=mode, engine=eng)
    for k, (b, cap) in enumerate(zip(blocks, caps)):
        rc, cons, prod, out = fn(b, cap)
        st, gc, gp, go = res[k]
        if mode == 1 and rc == 1 and st == 2:
            continue  # a full dst is reported distinctly in mode 1 (the reference's Buffer.t is unbounded)
        assert st == _ST[rc], (k, st, rc)
        assert (gp, go) == (prod, out), k
        if rc == 0:
            assert gc == cons, k


def test_minilzo_blocks_and_mutants_vs_oracle(eng):
    """fuzz/fuzz_lzo.ml:30-39: whatever minilzo compresses, Lzo.uncompress must decode identically; plus
    corrupted / truncated variants against the oracle in both error mappings."""
    sizes = [0, 1, 5, 100, 4096, 49152, 70000, 300000]
    plains = [workloads.text(n, 900 + n) for n in sizes]
    plains += [bytes(50000), (b"abcdefg" * 9000)[:60000], np.random.default_rng(3).integers(0, 256, 30000, dtype=np.uint8).tobytes()]
    comps = [workloads.lzo1x_1_compress(p) for p in plains]
    res = lzo.uncompress_batch(comps, [len(p) for p in plains], engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod) == (0, len(comps[k]), len(plains[k])), k
        assert out == plains[k], k
        rc, ref = workloads.lzo1x_decompress_safe(comps[k], len(plains[k]))
        assert rc == 0 and ref == out, k
    muts = _mutants([c for c in comps if len(c) > 10], 300, 11)
    for mode in (0, 1):
        _check(comps + muts, [1 << 19] * (len(comps) + len(muts)), mode, eng)
        _check(comps + muts, [40000] * (len(comps) + len(muts)), mode, eng)
        _check(comps, [max(0, len(p) - 1) for p in plains], mode, eng)


def test_c5_blocks_1mib(eng):
    """BASELINE config 5 at reduced count: 1 MiB blocks compressed by minilzo's lzo1x_1_compress."""
    comps, plains = workloads.c5_blocks(24)
    res = lzo.uncompress_batch(comps, [len(p) for p in plains], engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod) == (0, len(comps[k]), 1 << 20), k
        assert out == plains[k], k


def test_single_call_mirrors(eng):
    comp = workloads.lzo1x_1_compress(b"againandagainandagain" * 50)
    out = np.zeros(2000, dtype=np.uint8)
    r = lzo.uncompress(comp, out, engine=eng)
    assert r.is_ok() and r.value.tobytes() == b"againandagainandagain" * 50
    assert lzo.uncompress(comp, np.zeros(10, dtype=np.uint8), engine=eng) == Error("Invalid_argument")
    assert lzo.uncompress_with_buffer(comp, chunk=16, engine=eng) == Ok(b"againandagainandagain" * 50)
    assert lzo.uncompress_with_buffer(b"\x10\x00\x00", engine=eng) == Error("Malformed")


def test_cta_per_block_kernel_paths(eng):
    """Shapes that steer the CTA-per-block kernel (lzo_parallel.cu) through every branch: literal runs longer
    than its 16 KiB input window (moved by raw_copy), matches longer than a round's symbol budget (given to the
    warp-per-block kernel), rounds cut by the budget (highly compressible input), blocks spanning many
    rounds, odd source/destination alignment inside the slabs."""
    rng = np.random.default_rng(17)
    noise = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    txt = workloads.text(700000, 41)
    plains = [
        noise(20000), noise(300000),                           # one literal run each
        txt[:1000] + noise(70000) + txt[1000:3000],            # a long run between ordinary instructions
        bytes(200000), b"ab" * 150000, (b"0123456789abcdef" * 64 + b"x") * 300,   # long / periodic matches
        txt, txt[:333333] + bytes(5000) + txt[400000:500000],
        b"".join(txt[i * 3000:i * 3000 + 700] + noise(40) for i in range(150)),
    ]
    comps = [workloads.lzo1x_1_compress(p) for p in plains]
    for mode in (0, 1):
        res = lzo.uncompress_batch(comps, [len(p) for p in plains], mode=mode, engine=eng)
        for k, (st, cons, prod, out) in enumerate(res):
            assert (st, cons, prod) == (0, len(comps[k]), len(plains[k])), (mode, k)
            assert out == plains[k], (mode, k)
    # every alignment of the block inside the slabs (the input window and the drain are 16-byte machines)
    small = [workloads.lzo1x_1_compress(txt[o:o + 5000 + o]) for o in range(17)]
    res = lzo.uncompress_batch(small, [5000 + o for o in range(17)], engine=eng)
    for o, (st, cons, prod, out) in enumerate(res):
        assert (st, cons, prod, out) == (0, len(small[o]), 5000 + o, txt[o:o + 5000 + o]), o
    _check(comps, [len(p) - 1 for p in plains], 0, eng)
    _check(comps, [len(p) // 2 for p in plains], 1, eng)


def test_long_overlapping_copy_both_kernels(monkeypatch):
    """A self-overlapping match longer than the general kernel's 64 KiB ring minus its distance (ADVICE r1,
    lzo_decode.cu): hand-made blocks no compressor emits, through the default two-stage dispatch and through the
    general kernel alone (DCMP_INFLATE_MODE=serial), against the oracle."""
    lit = b"abcdefgh" + bytes(range(64, 128)) + workloads.text(30000, 3)
    blocks = [cases.lzo_block(lit[:max(off, 4 + off % 60)], off, ln) for off, ln in
              ((3, 70000), (5, 200000), (30000, 40000), (30000, 100000), (7, 65536 - 7 + 1), (4, 65536))]
    for serial in (False, True):
        if serial:
            monkeypatch.setenv("DCMP_INFLATE_MODE", "serial")
        e = _lib.Engine(0)
        for mode in (0, 1):
            _check(blocks, [300000] * len(blocks), mode, e)

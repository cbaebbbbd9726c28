"""The REAL kernel sources (decompress_b200/csrc/*.cu), compiled as C++ with -DDCMP_EMU and run on the CPU SIMT
emulator (tests/simt_emu, cooperative fibers: one per CUDA thread), against the CPU oracle.  No GPU needed: this
is what lets the GPU-less build box exercise the kernels' logic — rounds, checkpoints, head decodes, marker
replay, the LZO candidate/chase/replay pipeline, the two-stage dispatch — before a GPU minute is spent.  Test
infrastructure only; the product path is the CUDA build (tests/test_gpu_*.py)."""
import json
import os
import shutil
import zlib

import numpy as np

import pytest

import workloads
from decompress_b200 import batch
from tests import cases, oracle

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the emulator build needs g++")
_ST = {0: 0, 1: 16, 2: 17, 3: 18}  # oracle LZO code -> include/dcmp.h status


@pytest.fixture(scope="module")
def emu():
    from tests import emu as e
    e.build()
    return e


def _same(res, want):
    st, cons, prod, out = res
    rc, ocons, oprod, oout = want
    return st == rc and prod == oprod and out == oout and (rc != 0 or cons == ocons)


def test_inflate_valid_and_mutated_streams(emu):
    vs = cases.valid_streams()
    streams = [c for _, c, _ in vs]
    for caps in ([len(p) for _, _, p in vs], [max(0, len(p) - 1) for _, _, p in vs]):
        res, _ = emu.inflate_streams(streams, caps)
        for (name, c, _), r, cap in zip(vs, res, caps):
            assert _same(r, oracle.inflate_ns(c, cap)), name
    ms = cases.mutated_streams(60)
    res, retried = emu.inflate_streams([s for _, s in ms], [1 << 17] * len(ms))
    assert retried > 0  # the parallel kernel gives the odd ones to the general kernel
    for (name, c), r in zip(ms, res):
        assert _same(r, oracle.inflate_ns(c, 1 << 17)), name


def test_inflate_c2_streams_stored_and_multi_block(emu):
    comps, plains = workloads.c2_streams(3, workers=1)
    res, retried = emu.inflate_streams(comps, [65536] * 3, fmt=batch.FMT_ZLIB)
    assert retried == 0
    for r, c, p in zip(res, comps, plains):
        assert r == (0, len(c), 65536, p)
    txt = workloads.text(200000, 9)
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    multi = co.compress(txt) + co.flush()          # several dynamic blocks: headers parsed in line and ahead
    st0 = zlib.compressobj(0, zlib.DEFLATED, -15)
    stored = st0.compress(txt[:70001]) + st0.flush()  # stored blocks: the raw-copy path, odd sizes
    res, retried = emu.inflate_streams([multi, stored], [len(txt), 70001])
    assert retried == 0
    assert res[0] == (0, len(multi), len(txt), txt) and res[1] == (0, len(stored), 70001, txt[:70001])


def test_lzo_goldens_and_blocks(emu):
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lzo.json")))["vectors"]

    def src(v):
        if "src" in v:
            return bytes.fromhex(v["src"])
        pre, times, post = v["src_repeat"]
        return bytes.fromhex(pre) * times + bytes.fromhex(post)

    blocks = [src(v) for v in vec]
    # (which = 0 / 3: the CTA-per-block kernel in its small-round / large-round shape, the one a batch's size picks)
    for which in (0, 3):
        for mode, fn in ((0, oracle.lzo_uncompress), (1, oracle.lzo_uncompress_with_buffer)):
            res = emu.lzo_blocks(blocks, [1 << 16] * len(blocks), mode=mode, which=which)
            for v, b, r in zip(vec, blocks, res):
                rc, cons, prod, out = fn(b, 1 << 16)
                if mode == 1 and rc == 1 and r[0] == 2:
                    continue  # a full dst is reported distinctly in mode 1 (the reference's Buffer.t is unbounded)
                assert r[0] == _ST[rc] and r[2] == prod and r[3] == out, (v["name"], which)
    txt = workloads.text(120000, 4)
    plains = [txt, bytes(40000), os.urandom(30000), txt[:2000] + os.urandom(20000) + txt[2000:3000], b""]
    comps = [workloads.lzo1x_1_compress(p) for p in plains]
    for which in (0, 3):
        res = emu.lzo_blocks(comps, [len(p) for p in plains], which=which)
        for r, c, p in zip(res, comps, plains):
            assert r == (0, len(c), len(p), p), which


def test_lzo_long_overlapping_copy_general_kernel(emu):
    """A self-overlapping match longer than the ring minus its distance: its own ring writes wrap round onto the period
    it repeats (lzo_decode.cu read stale ring bytes there and still answered Ok)."""
    lit = b"abcdefgh" + bytes(range(64, 128)) + workloads.text(30000, 3)
    blocks = [cases.lzo_block(lit[:max(off, 4 + off % 60)], off, ln) for off, ln in
              ((3, 70000), (5, 200000), (30000, 40000), (30000, 100000), (7, 65536 - 7 + 1), (4, 65536))]
    caps = [300000] * len(blocks)
    for which in (2, 0, 3):
        res = emu.lzo_blocks(blocks, caps, mode=1, which=which)
        for b, r in zip(blocks, res):
            rc, cons, prod, out = oracle.lzo_uncompress_with_buffer(b, 300000)
            assert rc == 0 and r[0] == 0 and r[2] == prod and r[3] == out, (which, len(out))


def test_deflate_fixed_blocks_and_the_level_1_known_answer(emu):
    """test/test_ns.ml:1203-1222 (`encoder 1`): 56 zero bytes at level 1 are the five bytes 63 20 13 00 00 (a Fixed block)."""
    ins = [bytes(56), b"a", b"ab", b"abc", workloads.text(15, 1), workloads.text(16, 1), bytes(range(256))]
    res = emu.deflate_streams(ins, [len(x) + 64 for x in ins], level=1, fmt=batch.FMT_RAW)
    assert res[0][2].hex() == "6320130000"
    for x, (st, wr, comp) in zip(ins, res):
        assert st == 0 and zlib.decompress(comp, -15) == x and oracle.inflate_ns(comp, len(x))[:3] == (0, wr, len(x))
    assert all((res[k][2][0] >> 1) & 3 == 1 for k in range(6)) and (res[6][2][0] >> 1) & 3 == 0


def test_deflate_round_trip(emu):
    txt = workloads.text(150000, 2)
    ins = [txt, b"", txt[:100], bytes(70000), txt[:65537]]
    for level in (1, 6, 9):
        res = emu.deflate_streams(ins, [len(x) + len(x) // 8 + 64 for x in ins], level=level, fmt=batch.FMT_ZLIB)
        for (st, written, comp), x in zip(res, ins):
            assert st == 0 and zlib.decompress(comp) == x
            assert oracle.zl_inflate_ns(comp, len(x))[:3] == (0, written, len(x))


def test_deflate_window_edge_at_distance_32768_in_an_unaligned_batch(emu):
    """The far end of the window in streams that start at odd offsets of the batch: the compressor stages 16-byte granules,
    and the bytes around a half are window, not padding (tests/cases.py: deflate_window_edge_stream)."""
    edge = cases.deflate_window_edge_stream()
    ins = [b"abc", edge, b"z" * 7, edge[5:], edge]
    for level, align in ((1, 1), (6, 1), (9, 1), (6, 16)):   # (align = 1: streams at odd addresses of the slab)
        res = emu.deflate_streams(ins, [len(x) + len(x) // 8 + 64 for x in ins], level=level, fmt=batch.FMT_ZLIB, align=align)
        for (st, written, comp), x in zip(res, ins):
            assert st == 0 and zlib.decompress(comp) == x
            assert oracle.zl_inflate_ns(comp, len(x))[:3] == (0, written, len(x))


def test_deflate_random_batches_round_trip(emu):
    """Seeded random batches through both compressors: sizes around the 4 KiB / 32 KiB / 64 KiB boundaries, text, runs,
    noise, short periods, every level class, both slab alignments; zlib and the oracle's Inf must give the input back and
    consume exactly `written` bytes (fuzz/fuzz.ml:169-194)."""
    import random
    rng = random.Random(3)
    txt = bytes(workloads.text(150000, 77))

    def gen():
        k = rng.randrange(6)
        n = rng.choice([0, 1, 4, 5, 100, 4096, 4097, 32767, 32768, 32769, 65535, 65536, 65537, rng.randrange(1, 140000)])
        if k == 0: return txt[rng.randrange(1000):][:n]
        if k == 1: return bytes(n)
        if k == 2: return bytes(rng.randrange(256) for _ in range(min(n, 40000)))
        if k == 3: return (bytes(rng.randrange(256) for _ in range(rng.randrange(1, 300))) * (n // 2 + 1))[:n]
        if k == 4: return (txt[:rng.randrange(1, 40000)] * 8)[:n]
        return txt[:n // 2] + bytes(n - n // 2)

    for _ in range(12):
        ins = [gen() for _ in range(rng.randrange(1, 4))]
        level, align = rng.choice([1, 4, 6, 9]), rng.choice([1, 16])
        res = emu.deflate_streams(ins, [len(x) + len(x) // 8 + 128 for x in ins], level=level, fmt=batch.FMT_ZLIB, align=align)
        for (st, written, comp), x in zip(res, ins):
            assert st == 0 and zlib.decompress(comp) == x, (level, align, len(x))
            assert oracle.zl_inflate_ns(comp, len(x))[:3] == (0, written, len(x)), (level, align, len(x))


def test_fast_code_builder_is_complete_limited_and_as_cheap_as_t_make(emu):
    """df_build_code_fast (the CTA-per-chunk compressor's Huffman builder: sort, two-queue merge, pointer jumping, gen_bitlen's
    repair) against De.T.make as the oracle restates it: every used symbol has a code, no code is longer than the limit, the
    Kraft sum is exactly 1 (De.Inf.huffman rejects anything else, de.ml:544-550) and the block costs what zlib's tree costs —
    flat, skewed (the overflow repair), sparse and empty histograms of the three alphabets."""
    import random
    rng = random.Random(5)
    for n, mb in ((286, 15), (30, 15), (19, 7)):
        for t in range(60):
            kind = t % 6
            if kind == 0: f = [rng.choice([0, 0, 1, 2, 3, 50, 1000, rng.randrange(1 << 16)]) for _ in range(n)]
            elif kind == 1: f = [rng.randrange(1, 300) for _ in range(n)]
            elif kind == 2: f = [1 << (j % 17) for j in range(n)]
            elif kind == 3: f = [rng.randrange(200, 260) for _ in range(n)]
            elif kind == 4:
                f = [0] * n
                for _ in range(rng.randrange(0, 3)): f[rng.randrange(n)] = rng.randrange(1, 5)
            else: f = [rng.choice([0, 1]) for _ in range(n)]
            l1, c1, m1 = emu.t_make_fast(f, n, mb)
            l0, c0, m0 = oracle.t_make(f, n, mb)
            assert m1 == m0 and max(l1) <= mb and all(l1[i] > 0 for i in range(n) if f[i]), (n, t)
            assert sum(1 << (mb - l) for l in l1 if l) == 1 << mb, (n, t)
            assert sum(a * b for a, b in zip(f, l1)) <= sum(a * b for a, b in zip(f, l0)) * 1.002 + 2, (n, t)
            # canonical, bit-reversed codes: distinct and prefix-free follows from the Kraft equality + canonical order
            seen = {(l, c) for l, c in zip(l1, c1) if l}
            assert len(seen) == sum(1 for l in l1 if l), (n, t)


def test_encoder_kats_on_the_device_kernels(emu):
    """De.T.make and De.Def.encode over a Queue as the DEVICE builds them (deflate_tmake_kernel, deflate_queue_kernel), held to
    the reference's encoder known answers (test/test.ml:507-531, 1169-1237) and to the CPU oracle on random histograms."""
    k = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "encoder.json")))["kats"]
    st, out = emu.def_encode_queue(k["huffman_length_extra"]["cmds"], k["huffman_length_extra"]["last"])
    assert st == 0 and out.hex() == k["huffman_length_extra"]["bytes"]
    r = k["tree_rfc5322"]
    lens, codes, mc = emu.t_make(r["freqs"], 286)
    assert [lens[ord("6")], codes[ord("6")]] == r["literal_6"] and [lens[256], codes[256]] == r["eob"]
    assert (lens, codes, mc) == oracle.t_make(r["freqs"], 286)
    lens, codes, mc = emu.t_make([0] * 256 + [1], 286)
    assert [[n, l] for n, l in enumerate(lens) if l] == k["tree_0"]["lengths"]
    import random
    rng = random.Random(11)
    for length, maxl in ((286, 15), (30, 15), (19, 7)):
        for trial in range(12):
            f = [rng.choice([0, 0, 1, 2, 3, 50, 1000, rng.randrange(1 << 20)]) for _ in range(length)]
            if trial == 0:
                f = [0] * length
            if trial == 1:
                f = [0] * length; f[1] = 7
            if trial == 2:
                f = [1 << k2 for k2 in range(length)] if length <= 30 else [1 << (k2 % 22) for k2 in range(length)]  # skewed: overflow repair
            assert emu.t_make(f, length, maxl) == oracle.t_make(f, length, maxl), (length, trial)
    # a block of random commands: same bytes as the oracle's restatement of De.Def
    data, cmds = bytearray(), []
    for _ in range(1500):
        if data and rng.random() < 0.45:
            off = rng.randint(1, min(len(data), 32768)); ln = rng.randint(3, 258)
            for j in range(ln):
                data.append(data[-off])
            cmds.append(0x2000000 | ((ln - 3) << 16) | (off - 1))
        else:
            b = rng.choice(b"abcdefgh \n")
            data.append(b); cmds.append(b)
    cmds.append(256)
    st, out = emu.def_encode_queue(cmds, True)
    want, h = oracle.def_encode_dynamic(cmds, True)
    assert st == 0 and out == want and zlib.decompress(out, -15) == bytes(data)
    assert emu.def_encode_queue(cmds, True, cap=len(want) - 1)[0] == 2

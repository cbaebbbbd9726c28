"""GPU parity tests for the inflate path: the CUDA kernels, called through the C ABI
(libdcmp_b200.so), against the CPU oracle, the reference's golden vectors and system zlib.
Bit-exact: status, (consumed, produced) and every output byte."""
import json
import os
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import Error, Ok, _lib, batch, de, zl
from tests import cases, oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
BIG = 1 << 20


def _load(fn):
    with open(os.path.join(G, fn)) as f:
        return json.load(f)["vectors"]


@pytest.fixture(scope="module")
def eng():
    return _lib.Engine(0)


# ---- the reference's own KATs, written the way test/test_ns.ml writes them ---------------------
@pytest.mark.parametrize("v", _load("inflate_ns.json"), ids=lambda v: v["name"])
def test_ns_golden(v, eng):
    src = de.bigstring_of_string(bytes.fromhex(v["src"]))
    dst = de.bigstring_create(BIG if v["dst_len"] is None else v["dst_len"])
    res = de.Inf.Ns.inflate(src, dst, engine=eng)
    if "error" in v:
        assert res == Error(v["error"]), v["cite"]
    else:
        assert res == Ok(tuple(v["ok"])), v["cite"]
        if "out" in v:
            assert dst[:v["ok"][1]].tobytes() == bytes.fromhex(v["out"]), v["cite"]


def test_stream_goldens_in_one_batch(eng):
    vs = _load("inflate_stream.json")
    res = batch.inflate_streams([bytes.fromhex(v["src"]) for v in vs], [BIG] * len(vs), engine=eng)
    for v, (st, cons, prod, out) in zip(vs, res):
        if "malformed" in v:
            assert _lib.lib().dcmp_status_string(st).decode() == v["malformed"], v["cite"]
        else:
            assert st == 0 and out == bytes.fromhex(v["out"]), v["cite"]


# ---- differential against the oracle ------------------------------------------------------------
def _check_against_oracle(streams, caps, eng, fmt=batch.FMT_RAW):
    res = batch.inflate_streams(streams, caps, fmt=fmt, engine=eng)
    fn = oracle.inflate_ns if fmt == batch.FMT_RAW else oracle.zl_inflate_ns
    for k, (s, cap) in enumerate(zip(streams, caps)):
        rc, cons, prod, out = fn(s, cap)
        st, gcons, gprod, gout = res[k]
        assert st == rc, (k, _lib.STATUS_TAG.get(st), oracle.NS_ERRORS[rc])
        assert gprod == prod, k
        assert gout == out, k
        if rc == 0:
            assert gcons == cons, k


def test_valid_streams_vs_oracle_and_zlib(eng):
    vs = cases.valid_streams()
    streams = [c for _, c, _ in vs]
    caps = [len(p) + 64 for _, _, p in vs]
    res = batch.inflate_streams(streams, caps, engine=eng)
    for (name, comp, plain), (st, cons, prod, out) in zip(vs, res):
        assert st == 0, (name, _lib.STATUS_TAG.get(st))
        assert (cons, prod) == (len(comp), len(plain)), name
        assert out == plain, name
    _check_against_oracle(streams, caps, eng)
    # exact-size and one-short destinations (test_ns.ml:214-252)
    _check_against_oracle(streams, [len(p) for _, _, p in vs], eng)
    _check_against_oracle(streams, [max(0, len(p) - 1) for _, _, p in vs], eng)


def test_mutated_streams_vs_oracle(eng):
    ms = cases.mutated_streams(400)
    _check_against_oracle([s for _, s in ms], [1 << 18] * len(ms), eng)


def test_unaligned_offsets_and_ragged_batch(eng):
    """Streams packed back to back at odd offsets, empty streams in between."""
    vs = cases.valid_streams()[:8]
    streams = [b""] + [c for _, c, _ in vs] + [b"", b"\x03\x00"]
    caps = [0] + [len(p) for _, _, p in vs] + [5, 0]
    src = np.frombuffer(b"\xAA" + b"".join(streams), dtype=np.uint8).copy()
    off, so = [], 1
    for s in streams:
        off.append(so)
        so += len(s)
    dcap = np.array(caps, dtype=np.int64)
    doff = np.concatenate([[3], 3 + np.cumsum(dcap[:-1] + 1)])  # odd gaps
    dst = np.full(int(doff[-1] + dcap[-1] + 7), 0xEE, dtype=np.uint8)
    cons, prod, st = batch.inflate_host(batch.FMT_RAW, src, np.array(off), np.array([len(s) for s in streams]),
                                        dst, doff, dcap, engine=eng)
    for k, s in enumerate(streams):
        rc, c, p, out = oracle.inflate_ns(s, caps[k])
        assert (st[k], prod[k]) == (rc, p), k
        assert dst[doff[k]:doff[k] + p].tobytes() == out, k
        # nothing written outside what the stream produced (the reference never touches dst beyond o_pos)
        assert (dst[doff[k] + p:doff[k] + dcap[k] + 1] == 0xEE).all(), k
    assert (dst[:3] == 0xEE).all()


def test_c2_batch_256_streams(eng):
    """BASELINE config 2 at reduced count: independent 64 KiB single-dynamic-block zlib streams."""
    comps, plains = workloads.c2_streams(256, workers=4)
    res = batch.inflate_streams(comps, [65536] * len(comps), fmt=batch.FMT_ZLIB, engine=eng)
    for k, (st, cons, prod, out) in enumerate(res):
        assert st == 0 and cons == len(comps[k]) and prod == 65536, k
        assert out == plains[k], k
        assert zlib.adler32(out) == int.from_bytes(comps[k][-4:], "big")


def test_zlib_framing_errors(eng):
    z = zlib.compress(workloads.text(5000, 1))
    bad_sum = bytearray(z); bad_sum[-2] ^= 0x10
    bad_hdr = bytearray(z); bad_hdr[1] ^= 1
    streams = [z, bytes(bad_sum), bytes(bad_hdr), z[:1], z[:4], z[:-3], b"\x78\x9c\x03\x00\x00\x00\x00\x01"]
    _check_against_oracle(streams, [8192] * len(streams), eng, fmt=batch.FMT_ZLIB)
    dst = de.bigstring_create(8192)
    assert zl.Inf.Ns.inflate(z, dst, engine=eng) == Ok((len(z), 5000))
    assert zl.Inf.Ns.inflate(bytes(bad_sum), dst, engine=eng) == Error("Invalid_checksum")
    assert zl.Inf.Ns.inflate(bytes(bad_hdr), dst, engine=eng) == Error("Invalid_header")


def test_max_stored_block_c1(eng):
    """BASELINE config 1: a maximal 0xFFFF stored block, and 65536 bytes as two blocks (test.ml:1023-1035)."""
    payload = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8).tobytes()
    one = workloads.stored_block(payload[:65535])
    two = workloads.stored_block(payload[:65535], final=False) + workloads.stored_block(payload[65535:])
    res = batch.inflate_streams([one, two], [65535, 65536], engine=eng)
    assert res[0] == (0, len(one), 65535, payload[:65535])
    assert res[1] == (0, len(two), 65536, payload)


def test_device_resident_entry_point(eng):
    import torch
    comps, plains = workloads.c2_streams(64, workers=2)
    src, so, sl, dst, do, dc = batch.pack(comps, [65536] * 64)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    d_src, d_so, d_sl, d_dst, d_do, d_dc = t(src), t(so), t(sl), t(dst), t(do), t(dc)
    d_cons = torch.zeros(64, dtype=torch.int64, device=dev)
    d_prod = torch.zeros(64, dtype=torch.int64, device=dev)
    d_st = torch.full((64,), -1, dtype=torch.int32, device=dev)
    batch.inflate_device(batch.FMT_ZLIB, d_src, d_so, d_sl, d_dst, d_do, d_dc, d_cons, d_prod, d_st, engine=eng)
    torch.cuda.synchronize()
    assert d_st.cpu().tolist() == [0] * 64
    out = d_dst.cpu().numpy()
    for k in range(64):
        assert out[do[k]:do[k] + 65536].tobytes() == plains[k]
    assert eng.launch_count() > 0

"""Synthetic workloads for BASELINE.json's configs (shared by bench.py and tests/).

Nothing here is product code: it only *produces inputs* (with system zlib as the
compressor, which is the reference's own interop oracle, test/bin/simple.t) so that the
CUDA path, the CPU oracle and the benchmark all see the same seeded bytes.

C1  one stored block                       stored_block(payload)
C2  n x 64 KiB dynamic-Huffman zlib streams  c2_streams(n)
C3  enwik-style text for deflate            text(nbytes, seed)
C4  concatenated gzip members               c4_members(total, seed)
C5  LZO1X blocks (vendored minilzo)         c5_blocks(n, size, seed)
"""
import os
import struct
import zlib
from concurrent.futures import ProcessPoolExecutor

import numpy as np

_VOCAB = None


def _vocab():
    """3000 pseudo-words, Zipf-ranked; lengths 1..12 skewed short like English."""
    global _VOCAB
    if _VOCAB is None:
        rng = np.random.default_rng(12345)
        letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
        lp = 1.0 / np.arange(1, 27) ** 0.9
        lp /= lp.sum()
        nwords, maxlen = 3000, 12
        lens = np.clip(rng.poisson(4.2, nwords) + 1, 1, maxlen)
        lens[:60] = np.clip(lens[:60], 1, 4)  # frequent words are short
        mat = rng.choice(letters, size=(nwords, maxlen + 1), p=lp).astype(np.uint8)
        for i in range(nwords):
            mat[i, lens[i]] = 32  # trailing space
        mask = np.arange(maxlen + 1)[None, :] <= lens[:, None]
        p = 1.0 / np.arange(1, nwords + 1) ** 1.05
        p /= p.sum()
        _VOCAB = (mat, mask, np.cumsum(p))
    return _VOCAB


def text(nbytes, seed):
    """`nbytes` of word-soup text: Zipf vocabulary, sentence punctuation, line breaks."""
    if nbytes <= 0:
        return b""
    mat, mask, cdf = _vocab()
    rng = np.random.default_rng(seed)
    out = []
    have = 0
    while have < nbytes:
        nw = (nbytes - have) // 5 + 64
        idx = np.searchsorted(cdf, rng.random(nw))
        idx = np.minimum(idx, mat.shape[0] - 1)
        rows = mat[idx]
        m = mask[idx]
        # sprinkle punctuation / newlines on the trailing blank of some words
        u = rng.random(nw)
        last = m.sum(axis=1) - 1
        rows[np.arange(nw)[u < 0.06], last[u < 0.06]] = 10  # newline
        comma = (u >= 0.06) & (u < 0.10)
        rows[np.arange(nw)[comma], last[comma]] = 44
        # capitalise after newline-ish
        cap = np.roll(u < 0.06, 1)
        rows[cap, 0] = rows[cap, 0] - 32
        chunk = rows[m]
        out.append(chunk)
        have += chunk.size
    return np.concatenate(out)[:nbytes].tobytes()


def stored_block(payload, final=True):
    """RFC1951 stored block (BTYPE=00): de.ml:2566-2581 layout; payload <= 65535 bytes."""
    assert len(payload) <= 0xFFFF
    n = len(payload)
    return bytes([1 if final else 0]) + struct.pack("<HH", n, 0xFFFF - n) + bytes(payload)


def _c2_one(args):
    seed, size, fmt = args
    plain = text(size, seed)
    wbits = 15 if fmt == "zlib" else -15
    co = zlib.compressobj(6, zlib.DEFLATED, wbits, 9)  # memLevel 9: one block per 64 KiB
    comp = co.compress(plain) + co.flush()
    return comp, plain


def first_block_type(comp, fmt="zlib"):
    b = comp[2] if fmt == "zlib" else comp[0]
    return b & 1, (b >> 1) & 3  # (BFINAL, BTYPE)


def c2_streams(n, size=65536, seed0=0, fmt="zlib", workers=None, keep_plain=True):
    """n independent zlib streams, each ONE dynamic-Huffman block inflating to exactly `size` bytes.
    Returns (list of compressed bytes, list of plaintext bytes or None)."""
    workers = workers or min(32, os.cpu_count() or 1)
    jobs = [(seed0 + k, size, fmt) for k in range(n)]
    if workers > 1 and n >= 64:
        with ProcessPoolExecutor(workers) as ex:
            res = list(ex.map(_c2_one, jobs, chunksize=max(1, n // (workers * 4))))
    else:
        res = [_c2_one(j) for j in jobs]
    comps = [r[0] for r in res]
    for c in comps:
        assert first_block_type(c, fmt) == (1, 2), "C2 streams must be a single final dynamic block"
    return comps, ([r[1] for r in res] if keep_plain else None)


def gzip_member(plain, level=6):
    """One RFC1952 member: 10-byte header without optional fields, CRC-32 + ISIZE trailer."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9)
    body = co.compress(plain) + co.flush()
    hdr = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03"
    return hdr + body + struct.pack("<II", zlib.crc32(plain) & 0xFFFFFFFF, len(plain) & 0xFFFFFFFF)


def c4_members(total_plain, seed=0, lo=16 << 10, hi=256 << 10):
    """gzip members with log-uniform 16..256 KiB payloads until `total_plain` bytes; returns list of (member, plain)."""
    rng = np.random.default_rng(seed)
    out, have, k = [], 0, 0
    while have < total_plain:
        size = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        size = min(size, total_plain - have)
        plain = text(size, seed * 1000003 + k)
        out.append((gzip_member(plain), plain))
        have += size
        k += 1
    return out


def _c4_one(args):
    size, seed = args
    return gzip_member(text(size, seed))


def c4_blob(total_plain, seed=0, lo=16 << 10, hi=256 << 10, workers=None):
    """BASELINE config C4 at full size: gzip members with log-uniform 16..256 KiB payloads of the C3 text until
    `total_plain` bytes, generated by a process pool.  Returns (list of member bytes, int64 array of payload sizes);
    the plaintexts are not kept — a member's CRC-32 and ISIZE trailer is what the decoder is checked against."""
    rng = np.random.default_rng(seed)
    sizes, have = [], 0
    while have < total_plain:
        size = min(int(np.exp(rng.uniform(np.log(lo), np.log(hi)))), total_plain - have)
        sizes.append(size)
        have += size
    jobs = [(sz, seed * 1000003 + k) for k, sz in enumerate(sizes)]
    workers = workers or min(32, os.cpu_count() or 1)
    if workers > 1 and len(jobs) >= 64:
        with ProcessPoolExecutor(workers) as ex:
            members = list(ex.map(_c4_one, jobs, chunksize=max(1, len(jobs) // (workers * 8))))
    else:
        members = [_c4_one(j) for j in jobs]
    return members, np.array(sizes, dtype=np.int64)


_MINILZO = None


def minilzo():
    """The reference's own vendored LZO1X implementation (test/minilzo-2.10, built by oracle/Makefile into
    oracle/_ref/libminilzo.so — the reference's differential oracle for Lzo, fuzz/fuzz_lzo.ml)."""
    global _MINILZO
    if _MINILZO is None:
        import ctypes as C
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libminilzo.so")
        L = C.CDLL(path)
        for fn in (L.lzo1x_1_compress, L.lzo1x_decompress_safe):
            fn.restype = C.c_int
            fn.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.POINTER(C.c_ulong), C.c_void_p]
        _MINILZO = L
    return _MINILZO


def lzo1x_1_compress(data):
    import ctypes as C
    L = minilzo()
    data = bytes(data)
    out = C.create_string_buffer(len(data) + len(data) // 16 + 64 + 3)
    n = C.c_ulong(len(out))
    wrk = C.create_string_buffer(16384 * 8)
    rc = L.lzo1x_1_compress(data, len(data), out, C.byref(n), wrk)
    assert rc == 0
    return out.raw[:n.value]


def lzo1x_decompress_safe(comp, cap):
    import ctypes as C
    L = minilzo()
    out = C.create_string_buffer(max(cap, 1))
    n = C.c_ulong(cap)
    rc = L.lzo1x_decompress_safe(bytes(comp), len(comp), out, C.byref(n), None)
    return rc, out.raw[:n.value]


def c5_blocks(n, size=1 << 20, seed0=0):
    """n LZO1X blocks, each `size` bytes of the C3 text compressed by minilzo's lzo1x_1_compress.
    Returns (list of compressed blocks, list of plaintexts)."""
    plains = [text(size, 7000 + seed0 + k) for k in range(n)]
    return [lzo1x_1_compress(p) for p in plains], plains

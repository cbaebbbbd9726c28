"""Seeded DEFLATE test inputs shared by the CPU (oracle) and GPU (C-ABI) parity tests."""
import random
import zlib

import numpy as np

import workloads


def deflate_raw(data, level=6, mem=8, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategy)
    return co.compress(data) + co.flush()


def valid_streams():
    """(name, raw deflate stream, plaintext) covering stored / fixed / dynamic, multi-block, long runs, windows."""
    rng = np.random.default_rng(42)
    out = []
    txt = workloads.text(300000, 3)
    out.append(("empty", deflate_raw(b""), b""))
    out.append(("one_byte", deflate_raw(b"a"), b"a"))
    out.append(("text_64k_l6_m9", deflate_raw(txt[:65536], 6, 9), txt[:65536]))
    out.append(("text_300k_l6", deflate_raw(txt, 6, 8), txt))  # several dynamic blocks, 32 KiB window in use
    out.append(("text_300k_l1", deflate_raw(txt, 1, 8), txt))
    out.append(("text_300k_l9", deflate_raw(txt, 9, 9), txt))
    out.append(("text_fixed", deflate_raw(txt[:20000], 6, 8, zlib.Z_FIXED), txt[:20000]))
    out.append(("text_huffman_only", deflate_raw(txt[:50000], 6, 8, zlib.Z_HUFFMAN_ONLY), txt[:50000]))
    rnd = rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()
    out.append(("random_100k_stored", deflate_raw(rnd, 6), rnd))  # incompressible -> stored blocks
    out.append(("level0_stored_200k", deflate_raw(txt[:200000], 0), txt[:200000]))
    zeros = bytes(200000)
    out.append(("zeros_200k", deflate_raw(zeros, 9), zeros))  # dist=1 len=258 runs
    per = (b"abcdefghij" * 7 + b"\n") * 3000
    out.append(("periodic_71", deflate_raw(per, 6), per))
    rle = bytes(rng.integers(0, 4, 120000, dtype=np.uint8))
    out.append(("low_entropy", deflate_raw(rle, 6), rle))
    mixed = txt[:40000] + rnd[:40000] + zeros[:40000] + txt[40000:80000]
    out.append(("mixed_blocks", deflate_raw(mixed, 6), mixed))
    # sync-flush pieces: empty stored blocks between compressed blocks (de.ml flat + dynamic sequencing)
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    parts = b""
    for k in range(0, 90000, 7000):
        parts += co.compress(txt[k:k + 7000]) + co.flush(zlib.Z_SYNC_FLUSH)
    parts += co.flush()
    out.append(("sync_flush_pieces", parts, txt[:91000][:len(range(0, 90000, 7000)) * 7000]))
    far = txt[:1000] + bytes(rng.integers(0, 256, 31700, dtype=np.uint8)) + txt[:1000]
    out.append(("far_match_32k", deflate_raw(far, 9, 9), far))
    return out


def mutated_streams(count=300, seed=7):
    """Corrupted / truncated variants of valid streams: every reference error class shows up."""
    rnd = random.Random(seed)
    base = [s for _, s, _ in valid_streams() if len(s) < 60000]
    out = []
    for k in range(count):
        s = bytearray(rnd.choice(base))
        mode = rnd.randrange(5)
        if mode == 0 and len(s) > 2:      # truncate
            s = s[:rnd.randrange(1, len(s))]
        elif mode == 1:                    # flip a bit near the front (headers / tables)
            i = rnd.randrange(min(len(s), 80))
            s[i] ^= 1 << rnd.randrange(8)
        elif mode == 2:                    # flip a bit anywhere
            i = rnd.randrange(len(s))
            s[i] ^= 1 << rnd.randrange(8)
        elif mode == 3:                    # random garbage
            s = bytearray(rnd.randrange(256) for _ in range(rnd.randrange(1, 200)))
        else:                              # append garbage after a valid stream
            s += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 16)))
        out.append(("mut%03d_m%d" % (k, mode), bytes(s)))
    return out


def lzo_block(literals, off, length):
    """A hand-made LZO1X block (lzo.ml:315-393): an initial literal run, ONE match of `length` bytes at distance `off`
    (M3 for off <= 16384, M4 beyond), the end marker.  For the long self-overlapping copies no compressor emits."""
    assert len(literals) >= 4 and length >= 4 and 1 <= off <= min(49151, len(literals)) and off != 16384
    if len(literals) <= 238:
        out = bytearray([17 + len(literals)]) + bytearray(literals)
    else:  # opcode 0 in the main loop: a run of 18 + count() literals
        z, r = divmod(len(literals) - 18 - 1, 255)
        out = bytearray([0]) + bytes(z) + bytes([r + 1]) + bytearray(literals)
    ln = length - 2

    def count(base_bits, first):
        # len field: (first | len) when it fits the opcode's bits, else zeros * 255 + a last non-zero byte
        if ln <= base_bits:
            return bytes([first | ln])
        rest = ln - base_bits
        z, r = divmod(rest - 1, 255)
        return bytes([first]) + bytes(z) + bytes([r + 1])

    if off <= 16384:
        out += count(31, 0x20) + ((off - 1) << 2).to_bytes(2, "little")
    else:
        o = off - 16384
        out += count(7, 0x10 | ((o >> 11) & 8)) + ((o & 0x3fff) << 2).to_bytes(2, "little")
    return bytes(out + b"\x11\x00\x00")


def deflate_window_edge_stream():
    """A stream whose only good match at the start of every 32 KiB half is the start of the half before last... except that
    it is not: halves 1 and 2 begin with the same 16 bytes B, half 0 with 16 other bytes A whose first four hash (the
    compressor's hash4, 11 bits) like B's.  The candidate for the start of half 1 is therefore the start of half 0, exactly
    32768 back — the far end of the window — and the bytes there are A: a compressor that has let the staging of half 1
    spill over them (they are followed in the ring by half 2's first bytes, B) emits a match the decoder resolves to A."""
    def h(b4):
        return ((int.from_bytes(b4, "little") * 0x9E3779B1) & 0xFFFFFFFF) >> 21
    B = bytes(range(200, 216))
    A = None
    for x in range(1, 1 << 16):
        cand = bytes([x & 255, x >> 8, 7, 9])
        if h(cand) == h(B[:4]) and cand != B[:4]:
            A = cand + bytes(range(100, 112))
            break
    assert A is not None and h(B[:4]) != h(bytes(4))
    z = bytes(32768 - 16)
    return A + z + B + z + B + z[:20000]

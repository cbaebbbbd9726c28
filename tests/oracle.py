"""ctypes binding of oracle/liboracle.so — the CPU checker (TEST INFRASTRUCTURE ONLY).

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by
decompress_b200/.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

NS_ERRORS = ["Ok", "Unexpected_end_of_input", "Unexpected_end_of_output", "Invalid_kind_of_block",
             "Invalid_dictionary", "Invalid_complement_of_length", "Invalid_distance",
             "Invalid_distance_code", "Ref_exception", "Invalid_header", "Invalid_checksum"]
LZO_ERRORS = ["Ok", "Malformed", "Invalid_argument", "Invalid_dictionary"]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        sig = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        for name in ("orc_inflate_ns", "orc_zl_inflate_ns", "orc_lzo_uncompress", "orc_lzo_uncompress_with_buffer"):
            getattr(_lib, name).argtypes = sig
            getattr(_lib, name).restype = C.c_int
        for name in ("orc_adler32", "orc_crc32"):
            getattr(_lib, name).argtypes = [C.c_uint32, C.c_char_p, C.c_int64]
            getattr(_lib, name).restype = C.c_uint32
    return _lib


def _call(fn, src, cap):
    dst = C.create_string_buffer(max(cap, 1))
    cons, prod = C.c_int64(0), C.c_int64(0)
    rc = fn(bytes(src), len(src), dst, cap, C.byref(cons), C.byref(prod))
    return rc, cons.value, prod.value, dst.raw[:max(0, min(prod.value, cap))]


def inflate_ns(src, cap):
    """De.Inf.Ns.inflate: (status, consumed, produced, output-bytes)."""
    return _call(lib().orc_inflate_ns, src, cap)


def zl_inflate_ns(src, cap):
    return _call(lib().orc_zl_inflate_ns, src, cap)


def lzo_uncompress(src, cap):
    return _call(lib().orc_lzo_uncompress, src, cap)


def lzo_uncompress_with_buffer(src, cap):
    return _call(lib().orc_lzo_uncompress_with_buffer, src, cap)


def adler32(data, start=1):
    return lib().orc_adler32(start, bytes(data), len(data))


def crc32(data, start=0):
    return lib().orc_crc32(start, bytes(data), len(data))


def bench_batch(fmt, src, src_off, src_len, dst, dst_off, dst_cap, threads):
    """One timed pass of the CPU restatement over a batch (numpy arrays); returns (seconds, status, produced)."""
    import numpy as np

    L = lib()
    L.orc_bench_batch.restype = C.c_double
    L.orc_bench_batch.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    n = len(src_off)
    cons = np.zeros(n, dtype=np.int64)
    prod = np.zeros(n, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    secs = L.orc_bench_batch(fmt, n, p(src), p(src_off), p(src_len), p(dst), p(dst_off), p(dst_cap), p(cons), p(prod),
                             p(status), int(threads))
    return secs, status, prod


def t_make(freqs, length, max_length=15):
    """De.T.make ~length ?max_length freqs (oracle/deflate_t.c): (lengths[0..length), codes[0..length), max_code)."""
    import numpy as np

    L = lib()
    L.orc_t_make.restype = C.c_int
    L.orc_t_make.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4
    f = np.zeros(2 * 286 + 1, dtype=np.int32)
    f[:len(freqs)] = np.asarray(freqs, dtype=np.int32)[:2 * 286 + 1]
    lens = np.zeros(2 * 286 + 2, dtype=np.int32)
    codes = np.zeros(2 * 286 + 2, dtype=np.int32)
    blc = np.zeros(16, dtype=np.int32)
    mc = L.orc_t_make(length, max_length, C.c_void_p(f.ctypes.data), C.c_void_p(lens.ctypes.data), C.c_void_p(codes.ctypes.data),
                      C.c_void_p(blc.ctypes.data))
    return lens[:length].tolist(), codes[:length].tolist(), mc


def def_encode_dynamic(cmds, last=True):
    """One Dynamic block of De.Def over Queue commands (oracle/deflate_t.c): (bytes, (h_lit, h_dst, h_len))."""
    import numpy as np

    L = lib()
    L.orc_def_encode_dynamic.restype = C.c_int64
    L.orc_def_encode_dynamic.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    a = np.asarray(cmds, dtype=np.int32)
    out = np.zeros(8 * len(a) + 1024, dtype=np.uint8)
    h = np.zeros(3, dtype=np.int32)
    n = L.orc_def_encode_dynamic(C.c_void_p(a.ctypes.data), len(a), int(bool(last)), C.c_void_p(out.ctypes.data), len(out),
                                 C.c_void_p(h.ctypes.data))
    assert n >= 0
    return out[:n].tobytes(), tuple(h.tolist())

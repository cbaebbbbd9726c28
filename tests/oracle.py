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

"""Mirror of lib/de.mli for the GPU-backed path.

`Inf.Ns.inflate src dst` (de.mli:160-172) — one-shot inflate of a whole DEFLATE stream:
returns Ok((consumed, produced)) or Error(<constructor name>), the output is written into
`dst` (a writable buffer: bytearray / numpy uint8 array).
`Def.Ns.deflate ?level src dst` (de.mli:436-443) — one-shot deflate: Ok(len) or Error(...);
`Def.Ns.compress_bound len` (de.ml:3993-3996).
"""
import ctypes as C

import numpy as np

from . import _lib
from .result import Error, Ok

io_buffer_size = 65536  # de.ml:207


def _as_u8(buf, writable=False):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if writable and not a.flags.writeable:
        raise ValueError("dst must be writable")
    return a


def bigstring_create(n):
    return np.zeros(n, dtype=np.uint8)


def bigstring_of_string(s):
    return np.frombuffer(bytes(s), dtype=np.uint8)


class Inf:
    class Ns:
        errors = ("Unexpected_end_of_input", "Unexpected_end_of_output", "Invalid_kind_of_block",
                  "Invalid_dictionary", "Invalid_complement_of_length", "Invalid_distance",
                  "Invalid_distance_code")  # de.mli:151-157

        @staticmethod
        def inflate(src, dst, engine=None):
            eng = engine or _lib.default_engine()
            s = np.ascontiguousarray(_as_u8(src))
            d = _as_u8(dst, writable=True)
            cons, prod = C.c_int64(0), C.c_int64(0)
            st = _lib.lib().dcmp_inf_ns_inflate(eng.handle, C.c_void_p(s.ctypes.data if s.size else 0), s.size,
                                                C.c_void_p(d.ctypes.data if d.size else 0), d.size,
                                                C.byref(cons), C.byref(prod))
            if st < 0:
                _lib.check(st, "dcmp_inf_ns_inflate")
            if st == 0:
                return Ok((cons.value, prod.value))
            return Error(_lib.STATUS_TAG.get(st, st))


class Def:
    class Ns:
        errors = ("Invalid_compression_level", "Unexpected_end_of_output")  # de.mli:430-434

        @staticmethod
        def compress_bound(length):
            return int(_lib.lib().dcmp_deflate_bound(length, _lib.FMT_RAW))

        @staticmethod
        def deflate(src, dst, level=4, engine=None):
            eng = engine or _lib.default_engine()
            s = np.ascontiguousarray(_as_u8(src))
            d = _as_u8(dst, writable=True)
            wr = C.c_int64(0)
            st = _lib.lib().dcmp_def_ns_deflate(eng.handle, level, C.c_void_p(s.ctypes.data if s.size else 0), s.size,
                                                C.c_void_p(d.ctypes.data if d.size else 0), d.size, C.byref(wr))
            if st < 0:
                _lib.check(st, "dcmp_def_ns_deflate")
            if st == 0:
                return Ok(wr.value)
            return Error(_lib.STATUS_TAG.get(st, st))

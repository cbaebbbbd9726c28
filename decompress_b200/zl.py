"""Mirror of lib/zl.mli for the GPU-backed path: Zl.Inf.Ns.inflate (zl.ml:400-417), Zl.Def.Ns.deflate (zl.ml:610-630)."""
import ctypes as C

import numpy as np

from . import _lib
from .de import _as_u8
from .result import Error, Ok


class Inf:
    class Ns:
        @staticmethod
        def inflate(src, dst, engine=None):
            eng = engine or _lib.default_engine()
            s = np.ascontiguousarray(_as_u8(src))
            d = _as_u8(dst, writable=True)
            cons, prod = C.c_int64(0), C.c_int64(0)
            st = _lib.lib().dcmp_zl_inf_ns_inflate(eng.handle, C.c_void_p(s.ctypes.data if s.size else 0), s.size,
                                                   C.c_void_p(d.ctypes.data if d.size else 0), d.size,
                                                   C.byref(cons), C.byref(prod))
            if st < 0:
                _lib.check(st, "dcmp_zl_inf_ns_inflate")
            if st == 0:
                return Ok((cons.value, prod.value))
            return Error(_lib.STATUS_TAG.get(st, st))


class Def:
    class Ns:
        @staticmethod
        def compress_bound(length):
            return int(_lib.lib().dcmp_deflate_bound(length, _lib.FMT_ZLIB))

        @staticmethod
        def deflate(src, dst, level=4, engine=None):
            eng = engine or _lib.default_engine()
            s = np.ascontiguousarray(_as_u8(src))
            d = _as_u8(dst, writable=True)
            wr = C.c_int64(0)
            st = _lib.lib().dcmp_zl_def_ns_deflate(eng.handle, level, C.c_void_p(s.ctypes.data if s.size else 0), s.size,
                                                   C.c_void_p(d.ctypes.data if d.size else 0), d.size, C.byref(wr))
            if st < 0:
                _lib.check(st, "dcmp_zl_def_ns_deflate")
            if st == 0:
                return Ok(wr.value)
            return Error(_lib.STATUS_TAG.get(st, st))


class Higher:
    """Zl.Higher (zl.mli:265-368): compress ?level (default 6, zl.ml:634) and uncompress over the batch engine."""

    @staticmethod
    def compress(w, q, refill, flush, i, o, level=6, dynamic=True, engine=None):
        from .de import _higher_compress
        _higher_compress(_lib.FMT_ZLIB, level, refill, flush, i, o, engine)

    @staticmethod
    def uncompress(allocate, refill, flush, i, o, engine=None):
        from .de import _InfDecoder
        d = _InfDecoder("Manual", o, allocate(15) if allocate else None, engine, fmt=_lib.FMT_ZLIB)
        while True:
            r = d.decode()
            if r == "Await":
                n = refill(i)
                d.src(i, 0, n)
            elif r == "Flush":
                flush(o, len(o) - d.dst_rem())
                d.flush()
            elif r == "End":
                flush(o, len(o) - d.dst_rem())
                return Ok(None)
            else:
                return Error(("Msg", r[1]))

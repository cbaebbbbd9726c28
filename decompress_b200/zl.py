"""Mirror of lib/zl.mli for the GPU-backed path: Zl.Inf.Ns.inflate (zl.ml:400-417), Zl.Def.Ns.deflate (zl.ml:610-630)."""
import ctypes as C

import numpy as np

from . import _lib
from . import stream as _stream
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


def _inf_decoder(src, o, allocate=None, engine=None, api=None):
    """Zl.Inf.decoder src ~o ~allocate (zl.mli:54-64).  The reference threads a persistent value through
    decode / src / flush; here one mutable object plays every version of it.  `allocate` (the window) is accepted and
    ignored: the window is a device ring."""
    return _stream.InfDecoder(src, o, _lib.FMT_ZLIB, api=api, engine=engine)


def _def_encoder(q, w, level, src=None, dst=None, dynamic=True, engine=None, api=None):
    """Zl.Def.encoder ?dynamic ~q ~w ~level src dst (zl.mli:170-186); `q` / `w` are the reference's scratch for its serial
    matcher and are ignored.  src / dst: "Manual" (use .src / .dst), as the C ABI's dcmp_def_*."""
    return _stream.DefEncoder(_lib.FMT_ZLIB, level, api=api, engine=engine)


Inf.decoder = staticmethod(_inf_decoder)
Def.encoder = staticmethod(_def_encoder)


class Higher:
    """Zl.Higher (zl.mli:265-368): compress ?level (default 6, zl.ml:634) and uncompress, the reference's refill / flush
    loops over the streaming encoder / decoder."""

    @staticmethod
    def compress(w, q, refill, flush, i, o, level=6, dynamic=True, engine=None, api=None):
        _stream.compress(_lib.FMT_ZLIB, level, refill, flush, i, o, api=api, engine=engine)

    @staticmethod
    def uncompress(allocate, refill, flush, i, o, engine=None, api=None):
        err, _ = _stream.uncompress(_lib.FMT_ZLIB, refill, flush, i, o, api=api, engine=engine)
        return Ok(None) if err is None else Error(("Msg", err))

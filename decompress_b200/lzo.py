"""Mirror of lib/lzo.mli for the GPU-backed path (decode only; the compressor is out of scope).

    Lzo.uncompress input output            -> Ok(output[:len]) | Error(tag)      (lzo.mli:15-27)
    Lzo.uncompress_with_buffer ?chunk input -> Ok(bytes)        | Error(tag)      (lzo.mli:29-33)
"""
import ctypes as C

import numpy as np

from . import _lib
from .de import _as_u8
from .result import Error, Ok

_TAG = {16: "Malformed", 17: "Invalid_argument", 18: "Invalid_dictionary"}


def _batch1(eng, mode, s, d):
    n1 = np.array([s.size], dtype=np.int64)
    z = np.zeros(1, dtype=np.int64)
    cap = np.array([d.size], dtype=np.int64)
    cons = np.zeros(1, dtype=np.int64)
    prod = np.zeros(1, dtype=np.int64)
    st = np.zeros(1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data if a.size else 0)
    rc = _lib.lib().dcmp_lzo_uncompress_batch(eng.handle, mode, 1, p(s), p(z), p(n1), p(d), p(z), p(cap), p(cons), p(prod), p(st))
    _lib.check(rc, "dcmp_lzo_uncompress_batch")
    return int(st[0]), int(prod[0])


def uncompress(input, output, engine=None):
    """Lzo.uncompress (lib/lzo.ml:395-403): decodes into `output`, returns Ok(view of the produced prefix)."""
    eng = engine or _lib.default_engine()
    s = np.ascontiguousarray(_as_u8(input))
    d = _as_u8(output, writable=True)
    st, prod = _batch1(eng, 0, s, d)
    if st == 0:
        return Ok(d[:prod])
    return Error(_TAG.get(st, st))


def uncompress_with_buffer(input, chunk=0x1000, engine=None):
    """Lzo.uncompress_with_buffer (lib/lzo.ml:405-414): the reference grows a Buffer.t as it goes; here the
    output buffer is enlarged and the block decoded again whenever the device reports it full."""
    eng = engine or _lib.default_engine()
    s = np.ascontiguousarray(_as_u8(input))
    cap = max(int(chunk), 4 * s.size, 64)
    while True:
        d = np.zeros(cap, dtype=np.uint8)
        st, prod = _batch1(eng, 1, s, d)
        if st == 0:
            return Ok(d[:prod].tobytes())
        if st == 2:  # dst full: the reference's Buffer.t would have grown
            cap *= 4
            continue
        return Error(_TAG.get(st, st))


def uncompress_batch(blocks, caps, mode=0, engine=None):
    """n independent LZO1X blocks in one call: list of (status, consumed, produced, bytes)."""
    from . import batch

    eng = engine or _lib.default_engine()
    src, so, sl, dst, do, dc = batch.pack(blocks, caps)
    n = len(blocks)
    cons = np.zeros(n, dtype=np.int64)
    prod = np.zeros(n, dtype=np.int64)
    st = np.zeros(n, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    rc = _lib.lib().dcmp_lzo_uncompress_batch(eng.handle, mode, n, p(src), p(so), p(sl), p(dst), p(do), p(dc), p(cons),
                                              p(prod), p(st))
    _lib.check(rc, "dcmp_lzo_uncompress_batch")
    out = []
    for k in range(n):
        q = int(min(max(prod[k], 0), dc[k]))
        out.append((int(st[k]), int(cons[k]), int(prod[k]), dst[do[k]:do[k] + q].tobytes()))
    return out

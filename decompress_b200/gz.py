"""Mirror of lib/gz.mli's decode side for the GPU-backed path.

The reference's Gz.Inf is a streaming decoder that answers `End after ONE member (lib/gz.ml:344-356) and has no
one-shot (Ns) form; the batched equivalent here takes one gzip member per unit: header parsing (lib/gz.ml:463-492,
including the big-endian XLEN and the FHCRC coverage quirks), the DEFLATE body, then CRC-32 and ISIZE
(lib/gz.ml:344-356) — all on the device (csrc/gzip_frame.cu), so verified output never has to visit the host.
"""
import ctypes as C

import numpy as np

from . import _lib, batch
from .de import _as_u8
from .result import Error, Ok

FMT_GZIP = _lib.FMT_GZIP
_MSG = {1: "Unexpected end of input", 10: "Invalid checksum", 11: "Invalid GZip header", 12: "Invalid GZip header checksum",
        13: "Invalid input size"}


def inflate_members(members, caps, engine=None):
    """n gzip members in one call: list of (status, consumed, produced, bytes)."""
    return batch.inflate_streams(members, caps, fmt=FMT_GZIP, engine=engine)


def inflate_member(src, dst, engine=None):
    """One member: Ok((consumed, produced)) or Error(message prefix as in lib/gz.ml:285-296 / lib/de.ml:702-730)."""
    eng = engine or _lib.default_engine()
    s = np.ascontiguousarray(_as_u8(src))
    d = _as_u8(dst, writable=True)
    z = np.zeros(1, dtype=np.int64)
    sl = np.array([s.size], dtype=np.int64)
    dc = np.array([d.size], dtype=np.int64)
    cons, prod, st = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data if a.size else 0)
    rc = _lib.lib().dcmp_inflate_batch(eng.handle, FMT_GZIP, 1, p(s), p(z), p(sl), p(d), p(z), p(dc), p(cons), p(prod), p(st))
    _lib.check(rc, "dcmp_inflate_batch")
    if st[0] == 0:
        return Ok((int(cons[0]), int(prod[0])))
    return Error(_MSG.get(int(st[0]), _lib.lib().dcmp_status_string(int(st[0])).decode()))


def inflate_multi(buf, member_off, caps, engine=None):
    """Concatenated multi-member gzip (BASELINE config C4) with a sidecar index of member offsets: the reference
    cannot find member boundaries without decoding (lib/gz.ml:550, 582-610: the caller has to `reset` and re-feed
    `src_rem` bytes), so the batched form takes the boundaries as input and decodes every member in parallel."""
    buf = np.ascontiguousarray(_as_u8(buf))
    off = list(member_off) + [buf.size]
    members = [buf[off[k]:off[k + 1]] for k in range(len(member_off))]
    return inflate_members([m.tobytes() for m in members], caps, engine=engine)


def inflate_multi_sequential(buf, cap_per_member, engine=None):
    """The reference's way for a multi-member file without an index: decode a member, look at how much input it
    consumed, start the next one there (the `End -> src_rem -> reset loop of lib/gz.ml:550, 582-610)."""
    buf = np.ascontiguousarray(_as_u8(buf))
    pos, out = 0, []
    while pos < buf.size:
        st, cons, prod, data = inflate_members([buf[pos:].tobytes()], [cap_per_member], engine=engine)[0]
        if st != 0:
            return Error(_MSG.get(st, st)), out
        out.append(data)
        pos += cons
    return Ok(len(out)), out

"""Mirror of lib/gz.mli's decode side for the GPU-backed path.

The reference's Gz.Inf is a streaming decoder that answers `End after ONE member (lib/gz.ml:344-356) and has no
one-shot (Ns) form; the batched equivalent here takes one gzip member per unit: header parsing (lib/gz.ml:463-492,
including the big-endian XLEN and the FHCRC coverage quirks), the DEFLATE body, then CRC-32 and ISIZE
(lib/gz.ml:344-356) — all on the device (csrc/gzip_frame.cu), so verified output never has to visit the host.
"""
import ctypes as C

import numpy as np

from . import _lib, batch
from . import stream as _stream
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


def bgzf_index(buf):
    """Member boundaries of a BGZF file (bgzip / htslib) without decoding: every member carries its size in an FEXTRA
    sub-field (dcmp_gz_bgzf_index).  Returns (offsets, lengths, isizes) as int64 arrays, or None if `buf` is not BGZF."""
    b = np.ascontiguousarray(_as_u8(buf))
    L = _lib.lib()
    n = L.dcmp_gz_bgzf_index(C.c_void_p(b.ctypes.data if b.size else 0), b.size, None, None, None, 0)
    if n < 0:
        return None
    off, ln, isz = (np.zeros(n, dtype=np.int64) for _ in range(3))
    L.dcmp_gz_bgzf_index(C.c_void_p(b.ctypes.data if b.size else 0), b.size, C.c_void_p(off.ctypes.data), C.c_void_p(ln.ctypes.data),
                         C.c_void_p(isz.ctypes.data), n)
    return off, ln, isz


def inflate_bgzf(buf, engine=None):
    """A whole BGZF file in one batch: the index from the members' own size fields, every member an independent unit
    (CRC-32 / ISIZE checked on the device).  Returns Ok(bytes) or Error(message)."""
    b = np.ascontiguousarray(_as_u8(buf))
    idx = bgzf_index(b)
    if idx is None:
        return Error("Invalid GZip header")
    off, ln, isz = idx
    n = len(off)
    eng = engine or _lib.default_engine()
    dst_off = np.concatenate([[0], np.cumsum(isz)[:-1]]).astype(np.int64) if n else np.zeros(0, dtype=np.int64)
    dst = np.zeros(int(isz.sum()) + 16, dtype=np.uint8)
    cons, prod, st = batch.inflate_host(FMT_GZIP | _lib.GZIP_RFC_XLEN, b, off, ln, dst, dst_off, isz, engine=eng)
    bad = np.nonzero(st)[0]
    if len(bad):
        return Error(_MSG.get(int(st[bad[0]]), _lib.lib().dcmp_status_string(int(st[bad[0]])).decode()))
    return Ok(dst[:int(isz.sum())].tobytes())


def inflate_multi_sequential(buf, cap_per_member=None, engine=None, api=None, o_size=1 << 20):
    """The reference's way for a multi-member file without an index (lib/gz.ml:550, 582-610): Gz.Inf answers `End after
    ONE member; the caller looks at `src_rem` — how much of the input follows it —, resets the decoder and feeds the rest.
    One pass over the input with the streaming decoder: nothing is decoded or uploaded twice."""
    buf = np.ascontiguousarray(_as_u8(buf))
    o = np.zeros(o_size, dtype=np.uint8)
    d = _stream.InfDecoder("Manual", o, FMT_GZIP, api=api, engine=engine)
    pos, out, cur = 0, [], bytearray()
    step = 1 << 22
    fed_end = 0          # input handed over so far
    while True:
        r = d.decode()
        if r == "Await":
            n = min(step, buf.size - fed_end)
            d.src(buf, fed_end, n)   # n = 0: the end of the input
            fed_end += n
        elif r == "Flush":
            cur += o[:len(o) - d.dst_rem()].tobytes()
            d.flush()
        elif r == "End":
            out.append(bytes(cur))
            cur = bytearray()
            pos = fed_end - d.src_rem()   # where the next member starts
            if pos >= buf.size:
                return Ok(len(out)), out
            d.reset()
            fed_end = pos
        else:
            return Error(r[1].split(" (")[0]), out


class Inf:
    """Gz.Inf (gz.mli:56-149): decoder src ~o, decode, src, flush, reset, dst_rem, src_rem, write, filename, comment, os,
    extra ~key — mirrors of dcmp_inf_* with DCMP_FMT_GZIP."""

    @staticmethod
    def decoder(src, o, engine=None, api=None):
        return _stream.InfDecoder(src, o, FMT_GZIP, api=api, engine=engine)


OS = ("FAT", "Amiga", "VMS", "Unix", "VM", "Atari", "HPFS", "Macintosh", "Z", "CPM", "TOPS20", "NTFS", "QDOS", "Acorn")  # gz.ml:231-246


def os_of_int(v):
    return OS[v] if 0 <= v < len(OS) else "Unknown"


class Def:
    """Gz.Def (gz.mli:160-265): encoder src dst ?ascii ?hcrc ?filename ?comment ~mtime os ~q ~w ~level."""

    @staticmethod
    def encoder(mtime, os, level, ascii=False, hcrc=False, filename=None, comment=None, q=None, w=None, engine=None, api=None):
        os_v = OS.index(os) if isinstance(os, str) and os in OS else (255 if isinstance(os, str) else int(os))
        gzp = dict(mtime=mtime, os=os_v, ascii=ascii, hcrc=hcrc, filename=filename, comment=comment)
        return _stream.DefEncoder(FMT_GZIP, level, api=api, engine=engine, gz=gzp)


class Higher:
    """Gz.Higher (gz.mli:267-399): compress ?level (default 0... the reference's default is level 0, gz.ml:902-916) and
    uncompress, which returns the member's metadata."""

    @staticmethod
    def compress(w, q, refill, flush, time, cfg, i, o, level=0, filename=None, comment=None, engine=None, api=None):
        gzp = dict(mtime=int(time) & 0xFFFFFFFF, os=int(cfg.get("os", 255)) if isinstance(cfg, dict) else 255, ascii=False,
                   hcrc=False, filename=filename, comment=comment)
        _stream.compress(FMT_GZIP, level, refill, flush, i, o, api=api, engine=engine, gz=gzp)

    @staticmethod
    def uncompress(refill, flush, i, o, engine=None, api=None):
        err, d = _stream.uncompress(FMT_GZIP, refill, flush, i, o, api=api, engine=engine)
        if err is not None:
            return Error(("Msg", err))
        return Ok({"filename": d.filename(), "comment": d.comment(), "os": os_of_int(d.os())})

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


# ---- streaming facade ---------------------------------------------------------------------------------------------
class Window:
    """De.window (de.mli:59-75).  The GPU engine keeps its own sliding window; the value only carries `bits`."""

    def __init__(self, bits=15):
        self.bits = bits


def make_window(bits=15):
    return Window(bits)


class _InfDecoder:
    """De.Inf.decoder (de.mli:82-148) over the batch engine: the non-blocking `Await / `Flush / `End / `Malformed
    protocol is a host-side state machine (as in the reference, de.ml:651-753); the inflate itself runs on the GPU
    once the whole stream has been supplied (`src d buf off 0` signals end of input, de.ml:736-742).

    Signal sequence for a given output buffer `o`: `Await until the input is complete, then `Flush each time `o`
    is full (read o[0 .. len(o) - dst_rem()), call flush()), `End with the last partial buffer — the output side of
    what test/test.ml:1536-1568 pins.  On malformed input the bytes decoded before the error are flushed first,
    then `Malformed msg with the reference's message (de.ml:702-730)."""

    def __init__(self, src, o, w=None, engine=None, fmt=_lib.FMT_RAW):
        self._o = _as_u8(o, writable=True)
        self._eng = engine
        self._fmt = fmt
        self._manual = src == "Manual"
        self._chunks = []
        self._eof = not self._manual
        if not self._manual:
            kind, data = src  # ("String", bytes)
            assert kind == "String"
            self._chunks.append(bytes(data))
        self._out = None
        self._served = 0
        self._o_pos = 0
        self._status = 0
        self._consumed = 0
        self._done = False
        self._adler = 1

    # De.Inf.src (de.mli:118-124)
    def src(self, buf, off, length):
        if off < 0 or length < 0 or off + length > len(buf):
            raise ValueError("Invalid bounds (off: %d, len: %d)" % (off, length))  # de.ml:736-737
        if length == 0:
            self._eof = True
        else:
            self._chunks.append(bytes(buf[off:off + length]))

    def dst_rem(self):
        return len(self._o) - self._o_pos

    def src_rem(self):
        if self._out is None:
            return 0
        return max(0, sum(len(c) for c in self._chunks) - self._consumed)

    def flush(self):
        self._o_pos = 0

    def checksum(self):
        return self._adler

    def _run(self):
        import zlib as _z
        from . import batch
        data = b"".join(self._chunks)
        cap = max(4 * len(data), 1 << 16)
        while True:
            st, cons, prod, out = batch.inflate_streams([data], [cap], fmt=self._fmt, engine=self._eng)[0]
            if st == 2 and cap < (1 << 34):  # Unexpected_end_of_output: the provisional buffer was too small
                cap *= 4
                continue
            break
        self._status, self._consumed, self._out = st, cons, out
        self._adler = _z.adler32(out)

    def decode(self):
        if self._done:
            return "End" if self._status == 0 else ("Malformed", _lib.lib().dcmp_status_string(self._status).decode())
        if self._out is None:
            if not self._eof:
                return "Await"
            self._run()
        if self._o_pos == len(self._o) and self._served < len(self._out):
            return "Flush"  # idempotent until the caller flushes (zl.ml:288-291)
        n = min(len(self._o) - self._o_pos, len(self._out) - self._served)
        self._o[self._o_pos:self._o_pos + n] = np.frombuffer(self._out, dtype=np.uint8)[self._served:self._served + n]
        self._o_pos += n
        self._served += n
        if self._served < len(self._out):
            return "Flush"
        self._done = True
        if self._status == 0:
            return "End"
        return ("Malformed", _lib.lib().dcmp_status_string(self._status).decode())


def _decoder(src, o, w=None, engine=None):
    return _InfDecoder(src, o, w, engine)


Inf.decoder = staticmethod(_decoder)


class Higher:
    """De.Higher (de.mli:533-665): refill/flush callback loops over the streaming facade."""

    @staticmethod
    def uncompress(w, refill, flush, i, o, engine=None):
        d = Inf.decoder("Manual", o, w, engine=engine)
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

    @staticmethod
    def of_string(o, w, s, flush, engine=None):
        d = Inf.decoder(("String", s), o, w, engine=engine)
        while True:
            r = d.decode()
            if r == "Flush":
                flush(o, len(o) - d.dst_rem())
                d.flush()
            elif r == "End":
                flush(o, len(o) - d.dst_rem())
                return Ok(None)
            elif r == "Await":
                raise AssertionError("a `String source never awaits")
            else:
                return Error(("Msg", r[1]))


def _higher_compress(fmt, level, refill, flush, i, o, engine=None):
    """Shared body of De.Higher.compress / Zl.Higher.compress: pull the whole input through `refill`
    (it returns 0 at the end, de.ml:4517-4560), deflate it on the GPU, hand the result to `flush` in pieces of
    len(o)."""
    from . import batch
    chunks = []
    while True:
        n = refill(i)
        if n == 0:
            break
        chunks.append(bytes(i[:n]))
    data = b"".join(chunks)
    st, wr, comp = batch.deflate_streams([data], level=level, fmt=fmt, engine=engine)[0]
    if st != 0:
        raise RuntimeError(_lib.lib().dcmp_status_string(st).decode())
    ov = _as_u8(o, writable=True)
    for k in range(0, len(comp), len(ov)):
        n = min(len(ov), len(comp) - k)
        ov[:n] = np.frombuffer(comp, dtype=np.uint8)[k:k + n]
        flush(o, n)


def _compress(w, q, refill, flush, i, o, level=4, engine=None):
    """De.Higher.compress ~w ~q ~refill ~flush i o (de.mli:540-575).  `w` (Lz77 window) and `q` (Queue) are the
    reference's scratch for its serial matcher; the GPU path keeps its own (hash tables and token scratch on the
    device), so they are accepted and ignored."""
    _higher_compress(_lib.FMT_RAW, level, refill, flush, i, o, engine)


def _to_string(w, q, refill, i, buffer=4096, level=4, engine=None):
    """De.Higher.to_string ?buffer ~w ~q ~refill i (de.mli:640-665)."""
    out = bytearray()
    o = bigstring_create(max(int(buffer), 1))
    _higher_compress(_lib.FMT_RAW, level, refill, lambda buf, n: out.extend(buf[:n].tobytes()), i, o, engine)
    return bytes(out)


Higher.compress = staticmethod(_compress)
Higher.to_string = staticmethod(_to_string)

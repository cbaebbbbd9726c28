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
from . import stream as _stream
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


# ---- the command-level encoder: De.Queue commands, De.T, De.Def.encode over a Queue --------------------------------------
class Queue:
    """De.Queue's command encoding (de.mli:188-263, de.ml:2245-2266)."""
    EOB = 256

    @staticmethod
    def literal(chr_):
        return chr_ if isinstance(chr_, int) else ord(chr_)

    @staticmethod
    def copy(off, length):
        if not (1 <= off <= 32768 and 3 <= length <= 258):
            raise ValueError("Invalid copy (off: %d, len: %d)" % (off, length))
        return 0x2000000 | ((length - 3) << 16) | (off - 1)

    @staticmethod
    def of_list(cmds):
        """[("Literal", c) | ("Copy", off, len) | "End"] -> commands (Queue.of_list, de.ml:2321-2327)."""
        out = []
        for c in cmds:
            if c == "End":
                out.append(Queue.EOB)
            elif c[0] == "Literal":
                out.append(Queue.literal(c[1]))
            else:
                out.append(Queue.copy(c[1], c[2]))
        return out


class T:
    """De.T (de.mli:677-682): make ~length ?max_length freqs -> (lengths, codes, max_code), built on the device exactly as
    the reference builds it (zlib's tree)."""

    @staticmethod
    def make(length, freqs, max_length=15, engine=None):
        eng = engine or _lib.default_engine()
        f = np.zeros(length, dtype=np.int32)
        fr = np.asarray(freqs, dtype=np.int32)[:length]
        f[:len(fr)] = fr
        lens, codes, mc = np.zeros(length, dtype=np.int32), np.zeros(length, dtype=np.int32), C.c_int32(0)
        rc = _lib.lib().dcmp_t_make(eng.handle, length, max_length, C.c_void_p(f.ctypes.data), C.c_void_p(lens.ctypes.data),
                                    C.c_void_p(codes.ctypes.data), C.byref(mc))
        _lib.check(rc, "dcmp_t_make")
        return lens.tolist(), codes.tolist(), mc.value


def _encode_queue(cmds, last=True, cap=None, engine=None):
    """De.Def.encode over a Queue with a Dynamic block of its own frequencies (Def.dynamic_of_frequencies + encode,
    de.ml:2387-2407, 2583-2897), on the device: Ok(bytes) or Error(tag)."""
    eng = engine or _lib.default_engine()
    a = np.asarray(cmds, dtype=np.int32)
    cap = 8 * len(a) + 1024 if cap is None else cap
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    wr = C.c_int64(0)
    st = _lib.lib().dcmp_def_encode_queue(eng.handle, C.c_void_p(a.ctypes.data if a.size else 0), a.size, int(bool(last)),
                                          C.c_void_p(out.ctypes.data), cap, C.byref(wr))
    if st < 0:
        _lib.check(st, "dcmp_def_encode_queue")
    return Ok(out[:wr.value].tobytes()) if st == 0 else Error(_lib.STATUS_TAG.get(st, st))


Def.encode_queue = staticmethod(_encode_queue)


# ---- streaming facade ---------------------------------------------------------------------------------------------
class Window:
    """De.window (de.mli:59-75).  The GPU engine keeps its own sliding window; the value only carries `bits`."""

    def __init__(self, bits=15):
        self.bits = bits


def make_window(bits=15):
    return Window(bits)


class _InfDecoder(_stream.InfDecoder):
    """De.Inf.decoder (de.mli:82-148): the non-blocking `Await / `Flush / `End / `Malformed protocol, a host state
    machine behind the C ABI (dcmp_inf_*, csrc/stream.cu) over a resumable device kernel.  `decode` runs the kernel on
    the input handed over so far: output is flushed before the end of the input and memory is bounded (one undecodable
    tail of input, one launch's worth of output, the 32 KiB window on the device)."""

    def __init__(self, src, o, w=None, engine=None, fmt=_lib.FMT_RAW, api=None):
        super().__init__(src, o, fmt, api=api, engine=engine)


def _decoder(src, o, w=None, engine=None, api=None):
    return _InfDecoder(src, o, w, engine, api=api)


Inf.decoder = staticmethod(_decoder)


class Higher:
    """De.Higher (de.mli:533-665): the refill/flush callback loops over the streaming decoder / encoder."""

    @staticmethod
    def uncompress(w, refill, flush, i, o, engine=None, api=None):
        err, _ = _stream.uncompress(_lib.FMT_RAW, refill, flush, i, o, api=api, engine=engine)
        return Ok(None) if err is None else Error(("Msg", err))

    @staticmethod
    def of_string(o, w, s, flush, engine=None, api=None):
        d = Inf.decoder(("String", s), o, w, engine=engine, api=api)
        while True:
            r = d.decode()
            if r == "Flush":
                flush(o, len(o) - d.dst_rem())
                d.flush()
            elif r == "End":
                if len(o) - d.dst_rem() > 0:
                    flush(o, len(o) - d.dst_rem())
                return Ok(None)
            elif r == "Await":
                raise AssertionError("a `String source never awaits")
            else:
                return Error(("Msg", r[1]))


def _higher_compress(fmt, level, refill, flush, i, o, engine=None, api=None, gz=None):
    """Shared body of De.Higher.compress / Zl.Higher.compress / Gz.Higher.compress: `refill` until it returns 0
    (de.ml:4517-4560), segments deflated on the GPU as they fill up, `flush` for every full `o`."""
    _stream.compress(fmt, level, refill, flush, i, o, api=api, engine=engine, gz=gz)


def _compress(w, q, refill, flush, i, o, level=4, engine=None, api=None):
    """De.Higher.compress ~w ~q ~refill ~flush i o (de.mli:540-575).  `w` (Lz77 window) and `q` (Queue) are the
    reference's scratch for its serial matcher; the GPU path keeps its own (hash tables and token scratch on the
    device), so they are accepted and ignored."""
    _higher_compress(_lib.FMT_RAW, level, refill, flush, i, o, engine, api)


def _to_string(w, q, refill, i, buffer=4096, level=4, engine=None, api=None):
    """De.Higher.to_string ?buffer ~w ~q ~refill i (de.mli:640-665)."""
    out = bytearray()
    o = bigstring_create(max(int(buffer), 1))
    _higher_compress(_lib.FMT_RAW, level, refill, lambda buf, n: out.extend(buf[:n].tobytes()), i, o, engine, api)
    return bytes(out)


Higher.compress = staticmethod(_compress)
Higher.to_string = staticmethod(_to_string)

"""The streaming decoders and encoders of the reference — De.Inf.decoder (de.mli:82-148), Zl.Inf (zl.mli:30-124),
Gz.Inf (gz.mli:56-149), Zl.Def / Gz.Def encoders (zl.mli:141-263, gz.mli:160-265) — as thin ctypes mirrors of the C ABI's
dcmp_inf_* / dcmp_def_* (include/dcmp.h, csrc/stream.cu).  The state machines live behind the C ABI (C++); nothing is
decoded here.

`api` is the set of C entry points: the product library by default; the GPU-less tests pass the same state machines
built over the CPU SIMT emulator (tests/emu.py) to hold the protocol to the reference's sequences without a GPU."""
import ctypes as C

import numpy as np

from . import _lib

AWAIT, FLUSH, END, MALFORMED = 0, 1, 2, 3
_SIGNAL = {AWAIT: "Await", FLUSH: "Flush", END: "End"}


def _u8(buf, writable=False):
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if writable and not a.flags.writeable:
        raise ValueError("buffer must be writable")
    return a


class ProductApi:
    """dcmp_inf_* / dcmp_def_* of libdcmp_b200.so on one engine."""

    def __init__(self, engine=None):
        self.eng = engine or _lib.default_engine()
        self.L = _lib.lib()

    def inf_new(self, fmt, o_ptr, o_len):
        h = C.c_void_p()
        _lib.check(self.L.dcmp_inf_new(self.eng.handle, fmt, o_ptr, o_len, C.byref(h)), "dcmp_inf_new")
        return h

    def def_new(self, fmt, level):
        h = C.c_void_p()
        rc = self.L.dcmp_def_new(self.eng.handle, fmt, level, C.byref(h))
        if rc == -3:
            raise ValueError("Invalid level")  # De.Lz77.state (de.ml:4431-4437)
        _lib.check(rc, "dcmp_def_new")
        return h

    def __getattr__(self, name):  # inf_src -> dcmp_inf_src, gz_inf_os -> dcmp_gz_inf_os, ...
        return getattr(self.L, "dcmp_" + name)


class InfDecoder:
    """De.Inf.decoder / Zl.Inf.decoder / Gz.Inf.decoder.  src: "Manual" or ("String", bytes).  decode() answers "Await",
    "Flush", "End" or ("Malformed", message); see include/dcmp.h for the protocol."""

    def __init__(self, src, o, fmt=_lib.FMT_RAW, api=None, engine=None):
        self._api = api or ProductApi(engine)
        self._o = _u8(o, writable=True)
        self._h = self._api.inf_new(fmt, C.c_void_p(self._o.ctypes.data if self._o.size else 0), self._o.size)
        self._fmt = fmt
        self._keep = None          # the input handed over last: the C side reads it until the next `Await
        self._string = None
        if src != "Manual":
            kind, data = src
            assert kind == "String"
            self._string = bytes(data)
            self._feed(self._string)

    def __del__(self):
        try:
            if self._h:
                self._api.inf_free(self._h)
                self._h = None
        except Exception:
            pass

    def _feed(self, data):
        a = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
        self._keep = a
        rc = self._api.inf_src(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size)
        if rc != 0:
            raise ValueError("Invalid bounds")

    # De.Inf.src d buf off len (de.ml:733-742): len = 0 is the end of the input
    def src(self, buf, off, length):
        if off < 0 or length < 0 or off + length > len(buf):
            raise ValueError("Invalid bounds (off: %d, len: %d)" % (off, length))  # de.ml:736-737
        view = _u8(buf)[off:off + length]
        self._feed(np.ascontiguousarray(view))

    def decode(self):
        while True:
            r = self._api.inf_decode(self._h)
            if r < 0:
                _lib.check(r, "dcmp_inf_decode")
            if r == AWAIT and self._string is not None:
                self._feed(b"")  # a `String source is at its end when it is used up (de.ml:746)
                continue
            if r == MALFORMED:
                return ("Malformed", self._api.inf_error(self._h).decode())
            return _SIGNAL[r]

    def dst_rem(self):
        return int(self._api.inf_dst_rem(self._h))

    def src_rem(self):
        return int(self._api.inf_src_rem(self._h))

    def flush(self):
        self._api.inf_flush(self._h)

    def reset(self):
        self._api.inf_reset(self._h)
        if self._string is not None:
            self._feed(self._string)

    def checksum(self):
        return int(self._api.inf_checksum(self._h)) & 0xFFFFFFFF

    def write(self):
        return int(self._api.inf_write(self._h))

    def status(self):
        return int(self._api.inf_status(self._h))

    def launches(self):
        return int(self._api.inf_launches(self._h))

    # Gz.Inf.os / filename / comment / extra (gz.ml:612-633)
    def os(self):
        return int(self._api.gz_inf_os(self._h))

    def _field(self, which):
        p = C.c_void_p()
        n = self._api.gz_inf_field(self._h, which, C.byref(p))
        return None if n < 0 else C.string_at(p, n)

    def filename(self):
        return self._field(0)

    def comment(self):
        return self._field(1)

    def extra(self, key):
        if len(key) != 2:
            raise ValueError("Subfield ID must be 2 characters.")
        payload = self._field(2)
        if payload is None:
            return None
        kb = key.encode("latin1") if isinstance(key, str) else bytes(key)
        idx = 0
        while idx + 4 <= len(payload):
            k, ln = payload[idx:idx + 2], int.from_bytes(payload[idx + 2:idx + 4], "big")
            if idx + 4 + ln > len(payload):
                break
            if k == kb:
                return payload[idx + 4:idx + 4 + ln]
            idx += 4 + ln
        return None


class DefEncoder:
    """Zl.Def.encoder / Gz.Def.encoder (and, raw, De.Higher.compress's pair of Lz77 + Def).  encode() answers "Await",
    "Flush" or "End"."""

    def __init__(self, fmt, level, api=None, engine=None, gz=None):
        self._api = api or ProductApi(engine)
        self._h = self._api.def_new(fmt, level)
        self._keep_i = self._keep_o = None
        if gz is not None:
            fn, cm = gz.get("filename"), gz.get("comment")
            self._api.def_gz_header(self._h, int(gz.get("mtime", 0)) & 0xFFFFFFFF, int(gz.get("os", 255)), int(bool(gz.get("ascii"))),
                                    int(bool(gz.get("hcrc"))), fn, len(fn) if fn is not None else 0, cm, len(cm) if cm is not None else 0)

    def __del__(self):
        try:
            if self._h:
                self._api.def_free(self._h)
                self._h = None
        except Exception:
            pass

    def src(self, buf, off, length):
        if off < 0 or length < 0 or off + length > len(buf):
            raise ValueError("Invalid bounds (off: %d, len: %d)" % (off, length))
        a = np.ascontiguousarray(_u8(buf)[off:off + length])
        self._keep_i = a
        self._api.def_src(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size)

    def dst(self, buf, off, length):
        if off < 0 or length < 0 or off + length > len(buf):
            raise ValueError("Invalid bounds (off: %d, len: %d)" % (off, length))
        a = _u8(buf, writable=True)[off:off + length]
        self._keep_o = a
        self._api.def_dst(self._h, C.c_void_p(a.ctypes.data if a.size else 0), a.size)

    def encode(self):
        r = self._api.def_encode(self._h)
        if r < 0:
            _lib.check(r, "dcmp_def_encode")
        return _SIGNAL[r]

    def src_rem(self):
        return int(self._api.def_src_rem(self._h))

    def dst_rem(self):
        return int(self._api.def_dst_rem(self._h))


def uncompress(fmt, refill, flush, i, o, api=None, engine=None):
    """The loop of De.Higher.uncompress / Zl.Higher.uncompress / Gz.Higher.uncompress (de.ml:4562-4580, zl.ml:650-666):
    returns (None | message, decoder)."""
    d = InfDecoder("Manual", o, fmt, api=api, engine=engine)
    while True:
        r = d.decode()
        if r == "Await":
            n = refill(i)
            d.src(i, 0, n)
        elif r == "Flush":
            flush(o, len(o) - d.dst_rem())
            d.flush()
        elif r == "End":
            if len(o) - d.dst_rem() > 0:
                flush(o, len(o) - d.dst_rem())
            return None, d
        else:
            return r[1], d


def compress(fmt, level, refill, flush, i, o, api=None, engine=None, gz=None):
    """The loop of De.Higher.compress / Zl.Higher.compress / Gz.Higher.compress (de.ml:4517-4560, zl.ml:634-648)."""
    e = DefEncoder(fmt, level, api=api, engine=engine, gz=gz)
    e.dst(o, 0, len(o))
    while True:
        r = e.encode()
        if r == "Await":
            n = refill(i)
            e.src(i, 0, n)
        elif r == "Flush":
            flush(o, len(o) - e.dst_rem())
            e.dst(o, 0, len(o))
        else:
            return

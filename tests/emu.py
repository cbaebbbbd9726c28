"""ctypes binding of the CPU SIMT-emulated kernels (tests/_build/libemu_inflate.so).
TEST INFRASTRUCTURE ONLY — lets the kernel sources be debugged without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tests", "_build", "libemu_inflate.so")
_lib = None


def build():
    srcs = [os.path.join(ROOT, "tests", "simt_emu", f) for f in ("emu_inflate.cc", "simt_emu.cc", "simt_emu.h")]
    srcs += [os.path.join(ROOT, "decompress_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "decompress_b200", "csrc"))
             if f.endswith((".cu", ".cuh", ".inc", ".h"))]
    if os.path.exists(_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_SO) for s in srcs):
        return
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    extra = os.environ.get("DCMP_EMU_FLAGS", "").split()
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-DDCMP_EMU"] + extra + [
                           "-I" + os.path.join(ROOT, "tests", "simt_emu"), "-I" + os.path.join(ROOT, "decompress_b200", "csrc"),
                           "-o", _SO, srcs[0], srcs[1]])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.emu_inflate_batch.restype = C.c_int
        _lib.emu_inflate_batch.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    return _lib


def inflate_streams(streams, caps, fmt=0, mode=0, ctas=2):
    """Returns (list of (status, consumed, produced, bytes), number of streams the general kernel ran)."""
    from decompress_b200 import batch

    src, so, sl, dst, do, dc = batch.pack(streams, caps)
    n = len(streams)
    cons = np.zeros(n, dtype=np.int64)
    prod = np.zeros(n, dtype=np.int64)
    st = np.full(n, -1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    retried = lib().emu_inflate_batch(mode, fmt, n, p(src), p(so), p(sl), p(dst), p(do), p(dc), p(cons), p(prod), p(st), ctas)
    out = []
    for k in range(n):
        q = int(min(max(prod[k], 0), dc[k]))
        out.append((int(st[k]), int(cons[k]), int(prod[k]), dst[do[k]:do[k] + q].tobytes()))
    return out, retried


def lzo_blocks(blocks, caps, mode=0, which=0):
    """LZO kernels on the emulator: list of (status, consumed, produced, bytes).  which=0: CTA-per-block kernel, then the
    general one on the blocks it gave up (like engine.cu), small-round shape; which=3: the same with the large-round shape;
    which=2: the general kernel alone."""
    from decompress_b200 import batch

    src, so, sl, dst, do, dc = batch.pack(blocks, caps)
    n = len(blocks)
    cons = np.zeros(n, dtype=np.int64)
    prod = np.zeros(n, dtype=np.int64)
    st = np.full(n, -1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    L = lib()
    L.emu_lzo_batch2.restype = C.c_int
    L.emu_lzo_batch2.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 9
    L.emu_lzo_batch2(mode, which, n, p(src), p(so), p(sl), p(dst), p(do), p(dc), p(cons), p(prod), p(st))
    return [(int(st[k]), int(cons[k]), int(prod[k]), dst[do[k]:do[k] + int(min(max(prod[k], 0), dc[k]))].tobytes()) for k in range(n)]


def deflate_streams(inputs, caps, level=6, fmt=0, align=16):
    """Deflate kernels on the emulator: list of (status, written, bytes).  `align`: alignment of every stream inside the
    batch's slab (1: the streams follow each other byte for byte)."""
    from decompress_b200 import batch

    src, so, sl, dst, do, dc = batch.pack(inputs, caps, align=align)
    n = len(inputs)
    wr = np.zeros(n, dtype=np.int64)
    st = np.full(n, -1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    L = lib()
    L.emu_deflate_batch.restype = C.c_int
    L.emu_deflate_batch.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 8
    L.emu_deflate_batch(level, fmt, n, p(src), p(so), p(sl), p(dst), p(do), p(dc), p(wr), p(st))
    return [(int(st[k]), int(wr[k]), dst[do[k]:do[k] + int(wr[k])].tobytes()) for k in range(n)]


def t_make(freqs, length, max_length=15):
    """deflate_tmake_kernel on the emulator: (lengths, codes, max_code)."""
    L = lib()
    f = np.zeros(288, dtype=np.int32)
    fr = np.asarray(freqs, dtype=np.int32)[:length]
    f[:len(fr)] = fr
    lens, codes = np.zeros(288, dtype=np.int32), np.zeros(288, dtype=np.int32)
    L.emu_t_make.restype = C.c_int
    L.emu_t_make.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    mc = L.emu_t_make(length, max_length, C.c_void_p(f.ctypes.data), C.c_void_p(lens.ctypes.data), C.c_void_p(codes.ctypes.data))
    return lens[:length].tolist(), codes[:length].tolist(), mc


def t_make_fast(freqs, length, max_length=15):
    """The CTA-per-chunk compressor's builder (df_build_code_fast) on the emulator: (lengths, codes, max_code)."""
    L = lib()
    f = np.zeros(288, dtype=np.int32)
    fr = np.asarray(freqs, dtype=np.int32)[:length]
    f[:len(fr)] = fr
    lens, codes = np.zeros(288, dtype=np.int32), np.zeros(288, dtype=np.int32)
    L.emu_t_make_fast.restype = C.c_int
    L.emu_t_make_fast.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    mc = L.emu_t_make_fast(length, max_length, C.c_void_p(f.ctypes.data), C.c_void_p(lens.ctypes.data), C.c_void_p(codes.ctypes.data))
    return lens[:length].tolist(), codes[:length].tolist(), mc


def def_encode_queue(cmds, last=True, cap=None):
    """deflate_queue_kernel on the emulator: (status, bytes)."""
    L = lib()
    a = np.asarray(cmds, dtype=np.int32)
    cap = 8 * len(a) + 1024 if cap is None else cap
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    wr = C.c_int64(0)
    L.emu_def_encode_queue.restype = C.c_int
    L.emu_def_encode_queue.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    st = L.emu_def_encode_queue(C.c_void_p(a.ctypes.data), len(a), int(bool(last)), C.c_void_p(out.ctypes.data), cap, C.byref(wr))
    return st, out[:wr.value].tobytes()


class StreamApi:
    """The streaming state machines of the product (csrc/stream_core.h) instantiated over the emulator back ends
    (tests/simt_emu/emu_stream.h): same entry points as decompress_b200.stream.ProductApi, prefix emu_ instead of dcmp_."""
    _SIG = {
        "inf_new": (C.c_void_p, [C.c_int, C.c_void_p, C.c_int64]), "inf_free": (None, [C.c_void_p]),
        "inf_src": (None, [C.c_void_p, C.c_void_p, C.c_int64]), "inf_decode": (C.c_int, [C.c_void_p]),
        "inf_status": (C.c_int, [C.c_void_p]), "inf_error": (C.c_char_p, [C.c_void_p]), "inf_dst_rem": (C.c_int64, [C.c_void_p]),
        "inf_src_rem": (C.c_int64, [C.c_void_p]), "inf_flush": (None, [C.c_void_p]), "inf_reset": (None, [C.c_void_p]),
        "inf_checksum": (C.c_uint32, [C.c_void_p]), "inf_write": (C.c_int64, [C.c_void_p]), "inf_launches": (C.c_int64, [C.c_void_p]),
        "inf_buffered": (C.c_int64, [C.c_void_p]), "gz_inf_os": (C.c_int, [C.c_void_p]),
        "gz_inf_field": (C.c_int64, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
        "def_new": (C.c_void_p, [C.c_int, C.c_int, C.c_int64]), "def_free": (None, [C.c_void_p]),
        "def_gz_header": (None, [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]),
        "def_src": (None, [C.c_void_p, C.c_void_p, C.c_int64]), "def_dst": (None, [C.c_void_p, C.c_void_p, C.c_int64]),
        "def_encode": (C.c_int, [C.c_void_p]), "def_src_rem": (C.c_int64, [C.c_void_p]), "def_dst_rem": (C.c_int64, [C.c_void_p]),
    }

    def __init__(self, seg_bytes=1 << 18):
        self.L = lib()
        self.seg_bytes = seg_bytes
        for name, (res, args) in self._SIG.items():
            fn = getattr(self.L, "emu_" + name)
            fn.restype, fn.argtypes = res, args

    def inf_new(self, fmt, o_ptr, o_len):
        return C.c_void_p(self.L.emu_inf_new(fmt, o_ptr, o_len))

    def inf_src(self, h, p, n):
        self.L.emu_inf_src(h, p, n)
        return 0

    def def_new(self, fmt, level):
        if not 0 <= level <= 9:
            raise ValueError("Invalid level")
        return C.c_void_p(self.L.emu_def_new(fmt, level, self.seg_bytes))

    def __getattr__(self, name):
        return getattr(self.L, "emu_" + name)

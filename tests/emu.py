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
             if f.endswith((".cu", ".cuh"))]
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
    general one on the blocks it gave up (like engine.cu); which=2: the general kernel alone."""
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


def deflate_streams(inputs, caps, level=6, fmt=0):
    """Deflate kernels on the emulator: list of (status, written, bytes)."""
    from decompress_b200 import batch

    src, so, sl, dst, do, dc = batch.pack(inputs, caps)
    n = len(inputs)
    wr = np.zeros(n, dtype=np.int64)
    st = np.full(n, -1, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    L = lib()
    L.emu_deflate_batch.restype = C.c_int
    L.emu_deflate_batch.argtypes = [C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 8
    L.emu_deflate_batch(level, fmt, n, p(src), p(so), p(sl), p(dst), p(do), p(dc), p(wr), p(st))
    return [(int(st[k]), int(wr[k]), dst[do[k]:do[k] + int(wr[k])].tobytes()) for k in range(n)]

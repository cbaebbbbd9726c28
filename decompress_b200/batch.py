"""Batched entry points: n independent streams in one call (dcmp_inflate_batch*).

Host form: numpy uint8 slabs + int64 offset tables (copied to the device inside the call).
Device form: torch CUDA tensors (nothing copied; enqueued on torch's current stream).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FMT_RAW, FMT_ZLIB  # noqa: F401


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def pack(streams, caps, align=16):
    """Lay `streams` (bytes-like) out in one slab; returns (src, src_off, src_len, dst, dst_off, dst_cap)."""
    n = len(streams)
    src_len = np.array([len(s) for s in streams], dtype=np.int64)
    src_off = np.zeros(n, dtype=np.int64)
    dst_cap = np.array(caps, dtype=np.int64)
    dst_off = np.zeros(n, dtype=np.int64)
    so = do = 0
    for k in range(n):
        src_off[k] = so
        dst_off[k] = do
        so += (int(src_len[k]) + align - 1) // align * align
        do += (int(dst_cap[k]) + align - 1) // align * align
    src = np.zeros(max(so, 1), dtype=np.uint8)
    for k, s in enumerate(streams):
        src[src_off[k]:src_off[k] + len(s)] = np.frombuffer(bytes(s), dtype=np.uint8)
    dst = np.zeros(max(do, 1), dtype=np.uint8)
    return src, src_off, src_len, dst, dst_off, dst_cap


def inflate_host(fmt, src, src_off, src_len, dst, dst_off, dst_cap, engine=None):
    """dcmp_inflate_batch on host (numpy) buffers.  Returns (consumed, produced, status) int arrays."""
    eng = engine or _lib.default_engine()
    n = len(src_off)
    src_off = np.ascontiguousarray(src_off, dtype=np.int64)
    src_len = np.ascontiguousarray(src_len, dtype=np.int64)
    dst_off = np.ascontiguousarray(dst_off, dtype=np.int64)
    dst_cap = np.ascontiguousarray(dst_cap, dtype=np.int64)
    consumed = np.zeros(n, dtype=np.int64)
    produced = np.zeros(n, dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    rc = _lib.lib().dcmp_inflate_batch(eng.handle, fmt, n, _np_ptr(src), _np_ptr(src_off), _np_ptr(src_len),
                                       _np_ptr(dst), _np_ptr(dst_off), _np_ptr(dst_cap), _np_ptr(consumed),
                                       _np_ptr(produced), _np_ptr(status))
    _lib.check(rc, "dcmp_inflate_batch")
    return consumed, produced, status


def inflate_streams(streams, caps, fmt=FMT_RAW, engine=None):
    """Convenience: list of compressed byte strings -> list of (status, consumed, produced, output bytes)."""
    src, src_off, src_len, dst, dst_off, dst_cap = pack(streams, caps)
    consumed, produced, status = inflate_host(fmt, src, src_off, src_len, dst, dst_off, dst_cap, engine)
    out = []
    for k in range(len(streams)):
        p = int(min(max(produced[k], 0), dst_cap[k]))
        out.append((int(status[k]), int(consumed[k]), int(produced[k]), dst[dst_off[k]:dst_off[k] + p].tobytes()))
    return out


def inflate_device(fmt, d_src, d_src_off, d_src_len, d_dst, d_dst_off, d_dst_cap, d_consumed, d_produced, d_status,
                   engine=None, stream=None):
    """dcmp_inflate_batch_device on torch CUDA tensors; asynchronous on `stream` (default: torch's current)."""
    import torch

    eng = engine or _lib.default_engine(d_src.device.index or 0)
    if stream is None:
        stream = torch.cuda.current_stream(d_src.device).cuda_stream
    n = d_src_off.numel()
    rc = _lib.lib().dcmp_inflate_batch_device(
        eng.handle, fmt, n, C.c_void_p(d_src.data_ptr()), C.c_void_p(d_src_off.data_ptr()),
        C.c_void_p(d_src_len.data_ptr()), C.c_void_p(d_dst.data_ptr()), C.c_void_p(d_dst_off.data_ptr()),
        C.c_void_p(d_dst_cap.data_ptr()), C.c_void_p(d_consumed.data_ptr()), C.c_void_p(d_produced.data_ptr()),
        C.c_void_p(d_status.data_ptr()), C.c_void_p(stream))
    _lib.check(rc, "dcmp_inflate_batch_device")


def deflate_streams(inputs, caps=None, level=6, fmt=FMT_RAW, engine=None, align=16):
    """dcmp_deflate_batch on a list of byte strings: list of (status, written, compressed bytes).  `align`: alignment of
    every input and output window inside the two slabs (the C ABI takes any)."""
    eng = engine or _lib.default_engine()
    if caps is None:
        caps = [int(_lib.lib().dcmp_deflate_bound(len(x), fmt)) for x in inputs]
    src, so, sl, dst, do, dc = pack(inputs, caps, align=align)
    n = len(inputs)
    wr = np.zeros(n, dtype=np.int64)
    st = np.zeros(n, dtype=np.int32)
    rc = _lib.lib().dcmp_deflate_batch(eng.handle, level, fmt, n, _np_ptr(src), _np_ptr(so), _np_ptr(sl), _np_ptr(dst),
                                       _np_ptr(do), _np_ptr(dc), _np_ptr(wr), _np_ptr(st))
    _lib.check(rc, "dcmp_deflate_batch")
    return [(int(st[k]), int(wr[k]), dst[do[k]:do[k] + int(wr[k])].tobytes()) for k in range(n)]

"""ctypes binding of libdcmp_b200.so (include/dcmp.h).

The shared library is the product; this module only declares its symbols.  There is no
Python or CPU implementation behind it: if the library is missing, or there is no CUDA
device, the compute entry points raise.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCMP_LIB_PATH") or os.path.join(_HERE, "libdcmp_b200.so")  # override: tuning builds only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dcmp.h")

API_OK, API_NO_DEVICE, API_CUDA_ERROR, API_INVALID_ARGUMENT, API_OUT_OF_MEMORY = 0, -1, -2, -3, -4
FMT_RAW, FMT_ZLIB, FMT_GZIP = 0, 1, 2
GZIP_RFC_XLEN = 0x200  # OR-ed into FMT_GZIP: FEXTRA XLEN little-endian (RFC1952, BGZF) instead of the reference's big-endian

# De.Inf.Ns.error / Zl.Inf.Ns.error / Lzo.error constructors by status code (include/dcmp.h)
STATUS_TAG = {
    0: "Ok", 1: "Unexpected_end_of_input", 2: "Unexpected_end_of_output", 3: "Invalid_kind_of_block",
    4: "Invalid_dictionary", 5: "Invalid_complement_of_length", 6: "Invalid_distance",
    7: "Invalid_distance_code", 8: "Ref_exception", 9: "Invalid_header", 10: "Invalid_checksum",
    11: "Invalid_gzip_header", 12: "Invalid_gzip_header_checksum", 13: "Invalid_input_size",
    16: "Malformed", 17: "Invalid_argument", 18: "Invalid_dictionary", 32: "Invalid_compression_level",
}

_lib = None
p_u8 = C.c_void_p
p_i64 = C.c_void_p
p_i32 = C.c_void_p

_SIGNATURES = {
    "dcmp_abi_version": (C.c_int, []),
    "dcmp_device_count": (C.c_int, []),
    "dcmp_engine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "dcmp_engine_destroy": (None, [C.c_void_p]),
    "dcmp_last_error": (C.c_char_p, []),
    "dcmp_status_string": (C.c_char_p, [C.c_int]),
    "dcmp_engine_launch_count": (C.c_int64, [C.c_void_p]),
    "dcmp_engine_stats": (C.c_int, [C.c_void_p, p_i64, C.c_int]),
    "dcmp_inflate_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64, p_i64,
                                     p_i64, p_i64, p_i32]),
    "dcmp_inflate_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64, p_i64,
                                            p_i64, p_i64, p_i32, C.c_void_p]),
    "dcmp_inf_ns_inflate": (C.c_int, [C.c_void_p, p_u8, C.c_int64, p_u8, C.c_int64,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dcmp_zl_inf_ns_inflate": (C.c_int, [C.c_void_p, p_u8, C.c_int64, p_u8, C.c_int64,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dcmp_lzo_uncompress_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64, p_i64,
                                            p_i64, p_i64, p_i32]),
    "dcmp_lzo_uncompress_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64,
                                                   p_i64, p_i64, p_i64, p_i32, C.c_void_p]),
    "dcmp_lzo_uncompress": (C.c_int, [C.c_void_p, p_u8, C.c_int64, p_u8, C.c_int64, C.POINTER(C.c_int64)]),
    "dcmp_deflate_bound": (C.c_int64, [C.c_int64, C.c_int]),
    "dcmp_deflate_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64, p_i64,
                                     p_i64, p_i32]),
    "dcmp_deflate_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, p_u8, p_i64, p_i64, p_u8, p_i64,
                                            p_i64, p_i64, p_i32, C.c_void_p]),
    "dcmp_def_ns_deflate": (C.c_int, [C.c_void_p, C.c_int, p_u8, C.c_int64, p_u8, C.c_int64, C.POINTER(C.c_int64)]),
    "dcmp_zl_def_ns_deflate": (C.c_int, [C.c_void_p, C.c_int, p_u8, C.c_int64, p_u8, C.c_int64, C.POINTER(C.c_int64)]),
    "dcmp_engine_device": (C.c_int, [C.c_void_p]),
    "dcmp_gz_bgzf_index": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "dcmp_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "dcmp_host_unregister": (C.c_int, [C.c_void_p]),
    # streaming decoders / encoders (csrc/stream.cu)
    "dcmp_inf_new": (C.c_int, [C.c_void_p, C.c_int, p_u8, C.c_int64, C.POINTER(C.c_void_p)]),
    "dcmp_inf_free": (None, [C.c_void_p]),
    "dcmp_inf_src": (C.c_int, [C.c_void_p, p_u8, C.c_int64]),
    "dcmp_inf_dst": (C.c_int, [C.c_void_p, p_u8, C.c_int64]),
    "dcmp_inf_decode": (C.c_int, [C.c_void_p]),
    "dcmp_inf_status": (C.c_int, [C.c_void_p]),
    "dcmp_inf_error": (C.c_char_p, [C.c_void_p]),
    "dcmp_inf_dst_rem": (C.c_int64, [C.c_void_p]),
    "dcmp_inf_src_rem": (C.c_int64, [C.c_void_p]),
    "dcmp_inf_flush": (None, [C.c_void_p]),
    "dcmp_inf_reset": (None, [C.c_void_p]),
    "dcmp_inf_checksum": (C.c_uint32, [C.c_void_p]),
    "dcmp_inf_write": (C.c_int64, [C.c_void_p]),
    "dcmp_inf_launches": (C.c_int64, [C.c_void_p]),
    "dcmp_inf_buffered": (C.c_int64, [C.c_void_p]),
    "dcmp_gz_inf_os": (C.c_int, [C.c_void_p]),
    "dcmp_gz_inf_mtime": (C.c_uint32, [C.c_void_p]),
    "dcmp_gz_inf_field": (C.c_int64, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "dcmp_def_new": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "dcmp_def_free": (None, [C.c_void_p]),
    "dcmp_def_gz_header": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.c_char_p,
                                     C.c_int64]),
    "dcmp_def_src": (C.c_int, [C.c_void_p, p_u8, C.c_int64]),
    "dcmp_def_dst": (C.c_int, [C.c_void_p, p_u8, C.c_int64]),
    "dcmp_def_encode": (C.c_int, [C.c_void_p]),
    "dcmp_def_src_rem": (C.c_int64, [C.c_void_p]),
    "dcmp_def_dst_rem": (C.c_int64, [C.c_void_p]),
    "dcmp_def_encode_queue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, p_u8, C.c_int64, C.POINTER(C.c_int64)]),
    "dcmp_t_make": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcmp_adler32": (C.c_uint32, [C.c_uint32, p_u8, C.c_int64]),
    "dcmp_crc32": (C.c_uint32, [C.c_uint32, p_u8, C.c_int64]),
}


def declared_symbols():
    """Every function name include/dcmp.h declares (used by the CPU test that checks the exports)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcmp_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libdcmp_b200.so is not built (%s). Build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
                "or `make -C decompress_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def last_error():
    return lib().dcmp_last_error().decode()


class DcmpError(RuntimeError):
    pass


def check(rc, what):
    if rc == API_OK:
        return
    names = {API_NO_DEVICE: "no CUDA device", API_CUDA_ERROR: "CUDA error", API_INVALID_ARGUMENT: "invalid argument",
             API_OUT_OF_MEMORY: "out of device memory"}
    raise DcmpError("%s failed: %s [%s]" % (what, names.get(rc, rc), last_error()))


class Engine:
    """One CUDA device + stream + scratch (dcmp_engine).  Single owner, like a De.Inf.decoder."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().dcmp_engine_create(device, C.byref(h)), "dcmp_engine_create")
        self._h = h
        self.device = device

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return lib().dcmp_engine_launch_count(self._h)

    def stats(self):
        """dcmp_engine_stats: units submitted to / redone by the general kernels of the two-stage dispatch."""
        out = (C.c_int64 * 5)()
        check(lib().dcmp_engine_stats(self._h, out, 5), "dcmp_engine_stats")
        return dict(zip(("inflate_streams", "inflate_redone", "lzo_blocks", "lzo_redone", "launches"), list(out)))

    def close(self):
        if self._h:
            lib().dcmp_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_engine(device=0):
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]

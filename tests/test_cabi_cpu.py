"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/dcmp.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import os
import ctypes as C

import pytest

from decompress_b200 import _lib


def test_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), "libdcmp_b200.so does not export %s" % name
        assert name in _lib._SIGNATURES, "decompress_b200/_lib.py does not bind %s" % name


def test_abi_version_and_strings():
    lib = _lib.lib()
    assert lib.dcmp_abi_version() == 2
    # de.ml:1557-1567
    want = {1: "Unexpected end of input", 2: "Unexpected end of output", 3: "Invalid kind of block",
            4: "Invalid dictionary", 5: "Invalid complement of length", 6: "Invalid distance",
            7: "Invalid distance code"}
    for code, msg in want.items():
        assert lib.dcmp_status_string(code).decode() == msg


def test_host_checksums_match_libz():
    import zlib
    lib = _lib.lib()
    data = bytes(range(256)) * 300 + b"tail"
    assert lib.dcmp_adler32(1, data, len(data)) == zlib.adler32(data)
    assert lib.dcmp_crc32(0, data, len(data)) == zlib.crc32(data)
    assert lib.dcmp_adler32(1, b"", 0) == 1
    # gzip golden: CRC-32 of "foo" (test/test.ml:1703-1704)
    assert lib.dcmp_crc32(0, b"foo", 3) == 0x8C736521


def test_no_device_is_a_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.lib().dcmp_device_count() == 0
    with pytest.raises(_lib.DcmpError, match="no CUDA device"):
        _lib.Engine(0)
    h = C.c_void_p()
    assert _lib.lib().dcmp_engine_create(0, C.byref(h)) == _lib.API_NO_DEVICE


def test_ocaml_shim_is_consistent_with_the_c_abi():
    """ocaml/ cannot be compiled here (no OCaml toolchain): hold it to the header at least — every dcmp_* call of the
    stubs is a declared entry point, every `external` of dcmp.ml has its CAMLprim (native and bytecode names)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stubs = open(os.path.join(root, "ocaml", "dcmp_stubs.c")).read()
    ml = open(os.path.join(root, "ocaml", "dcmp.ml")).read()
    declared = set(_lib.declared_symbols())
    called = set(re.findall(r"\b(dcmp_(?!stub_)[a-z0-9_]+)\s*\(", stubs))
    assert called <= declared, sorted(called - declared)
    prims = set(re.findall(r"CAMLprim value (dcmp_stub_[a-z0-9_]+)", stubs))
    wanted = set(re.findall(r'"(dcmp_stub_[a-z0-9_]+)"', ml))
    assert wanted <= prims, sorted(wanted - prims)
    for f in ("de.ml", "zl.ml", "gz.ml", "lzo.ml", "dune"):
        assert os.path.exists(os.path.join(root, "ocaml", f))

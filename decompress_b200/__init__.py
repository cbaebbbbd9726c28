"""decompress_b200 — B200-native DEFLATE / zlib / gzip / LZO1X engine behind mirage/decompress's
module signatures.  `de`, `zl`, `gz`, `lzo` mirror lib/de.mli, lib/zl.mli, lib/gz.mli, lib/lzo.mli;
`batch` is the batched (many independent streams) form the GPU wants.  All compute goes through
libdcmp_b200.so (include/dcmp.h); there is no CPU implementation in this package."""
from . import _lib  # noqa: F401
from .result import Ok, Error  # noqa: F401
from . import batch, de, zl, gz, lzo  # noqa: F401

__all__ = ["Ok", "Error", "batch", "de", "zl", "gz", "lzo"]

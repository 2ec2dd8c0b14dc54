"""Importable alias of the product package.

The package directory is named after the reference crate (`rust-lz-fear_b200/`), which is not a
valid Python identifier; this stub makes it importable as `lz_fear_b200` by pointing its
`__path__` at that directory.  All code lives there.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rust-lz-fear_b200")
__path__.insert(0, _REAL)

from ._native import (  # noqa: E402,F401
    Context, NativeLibraryError, LzfCallError, load_library, library_path,
)
from . import raw, framed  # noqa: E402,F401
from .framed import (  # noqa: E402,F401
    CompressionSettings, CompressionError, LZ4FrameReader, LZ4FrameIoReader, DecompressionError,
    decompress_frame, MAGIC, WINDOW_SIZE,
    # CompressionError / DecompressionError variants (src/framed/compress.rs:33-42, decompress.rs:20-44)
    ReadError, WriteError, InvalidBlockSize, InputError, CodecError, HeaderParseError, WrongMagic, HeaderChecksumFail,
    BlockChecksumFail, FrameChecksumFail, BlockLengthOverflow, BlockSizeOverflow,
)

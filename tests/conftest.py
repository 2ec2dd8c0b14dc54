import ctypes
import json
import lzma
import os
import subprocess
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SIMT_LIB = os.path.join(ROOT, "tests", "simt", "libsimt_lzfear.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def vectors():
    return json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))


@pytest.fixture(scope="session")
def issue15_input():
    return lzma.decompress(open(os.path.join(GOLDEN, "issue15_input.bin.xz"), "rb").read())


@pytest.fixture(scope="session")
def corpora():
    """{'decode': [(name, bytes)...], 'interop_decode': [...], 'roundtrip_fuzz': [...]}"""
    out = {"decode": [], "interop_decode": [], "roundtrip_fuzz": []}
    with tarfile.open(os.path.join(GOLDEN, "corpus.tar.xz")) as tf:
        for m in tf.getmembers():
            if m.isfile():
                d, name = m.name.split("/", 1)
                out[d].append((name, tf.extractfile(m).read()))
    for v in out.values():
        v.sort()
    return out


@pytest.fixture(scope="session")
def liblz4():
    try:
        lib = ctypes.CDLL("liblz4.so.1")
    except OSError:
        pytest.skip("liblz4.so.1 not available")
    lib.LZ4_compress_default.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    lib.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    lib.LZ4_compressBound.argtypes = [ctypes.c_int]
    return lib


@pytest.fixture(scope="session")
def simt_lib_path():
    """Builds the CPU SIMT-emulator build of the product sources (tests/simt, test infrastructure)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "simt")])
    return SIMT_LIB


class _Backend:
    """A Context of the product binding bound to one build of the C ABI."""

    def __init__(self, name, ctx, scale):
        self.name, self.ctx, self.scale = name, ctx, scale

    def fresh(self):
        """A second Context of the same build, created NOW: lzf_create reads the LZF_B200_* tuning knobs once, so a test
        that sets one with monkeypatch.setenv takes its context from here (use as a context manager)."""
        import contextlib
        from lz_fear_b200 import _native

        @contextlib.contextmanager
        def cm():
            ctx = _native.Context(0)
            try:
                yield _Backend(self.name, ctx, self.scale)
            finally:
                ctx.close()
        return cm()


@pytest.fixture(scope="session")
def emu(simt_lib_path):
    """The product Python binding pointed at the SIMT-emulated build: exercises the REAL kernel and
    C-ABI sources on the CPU (never used outside tests)."""
    from lz_fear_b200 import _native
    saved = (_native._lib, _native._lib_path)
    _native.load_library(simt_lib_path)
    ctx = _native.Context(0)
    yield _Backend("simt-emu", ctx, 1)
    ctx.close()
    _native._lib, _native._lib_path = saved


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.join(ROOT, "rust-lz-fear_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("lzf_build", os.path.join(ROOT, "rust-lz-fear_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from lz_fear_b200 import _native
    _native._lib = None
    _native._lib_path = None
    _native.load_library()
    ctx = _native.Context(0)
    yield _Backend("cuda", ctx, 16)
    ctx.close()


def rng_bytes(n, seed):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()

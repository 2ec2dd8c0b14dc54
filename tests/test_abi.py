"""The shipped C-ABI library: builds for sm_100a, loads, exports every symbol include/*.h declares,
and refuses to work without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import importlib.util
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    spec = importlib.util.spec_from_file_location("lzf_build", os.path.join(ROOT, "rust-lz-fear_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lzfear_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lzf_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = _build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "liblzfear_b200.so does not export %s" % name
    lib.lzf_abi_version.restype = ctypes.c_int
    assert lib.lzf_abi_version() == 3


def test_python_binding_covers_the_header():
    from lz_fear_b200 import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == declared_symbols()


def test_library_targets_sm_100a_and_contains_the_kernels():
    lib_path = _build()
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    syms = subprocess.run(["cuobjdump", "-res-usage", lib_path], capture_output=True, text=True).stdout
    for kern in ("encode_blocks_kernel", "decode_blocks_kernel", "xxh32_ranges_kernel", "frame_layout_kernel",
                 "frame_assemble_kernel", "frame_walk_kernel"):
        assert kern in syms, kern


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from lz_fear_b200 import _native
    saved = (_native._lib, _native._lib_path)
    try:
        _native._lib = None
        _native.load_library(_build())
        with pytest.raises(_native.NativeLibraryError):
            _native.Context(0)
    finally:
        _native._lib, _native._lib_path = saved


def test_missing_library_fails_loudly(tmp_path):
    from lz_fear_b200 import _native
    with pytest.raises(_native.NativeLibraryError):
        _native.load_library(str(tmp_path / "nope.so"))


def test_product_sources_never_touch_the_oracle_or_the_emulator():
    pkg = os.path.join(ROOT, "rust-lz-fear_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "lzf_oracle" not in txt, f
                assert "simt_emu.h\"" not in txt and "libsimt" not in txt, f


def _rust_externs(path):
    """{name: [rust param types], ...} of the `extern "C"` block of a Rust source file."""
    import re
    txt = open(path).read()
    block = txt[txt.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = [p.split(":", 1)[1].strip() for p in m.group(2).split(",") if ":" in p]
        out[m.group(1)] = (params, (m.group(3) or "").strip())
    return out


def test_rust_sys_crate_matches_the_header():
    """rust/lz-fear-b200-sys/src/lib.rs declares every symbol of include/lzfear_b200.h with the header's parameter
    list (count, pointer depth, constness, integer width), and is exactly what rust/gen_sys.py generates."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_sys", os.path.join(ROOT, "rust", "gen_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    lib_rs = os.path.join(ROOT, "rust", "lz-fear-b200-sys", "src", "lib.rs")
    assert open(lib_rs).read() == gen.generate(), "lib.rs is stale: run python rust/gen_sys.py"
    rust = _rust_externs(lib_rs)
    protos = gen.c_prototypes()
    assert sorted(rust) == sorted(p[0] for p in protos) == sorted(declared_symbols())
    width = {"u8": "uint8_t", "u32": "uint32_t", "u64": "uint64_t", "i32": "int32_t", "usize": "size_t", "c_int": "int",
             "c_void": "void", "c_char": "char"}
    for name, ret, params in protos:
        rparams, rret = rust[name]
        assert len(rparams) == len(params), name
        for (ctype, _pn), rt in zip(params, rparams):
            assert rt.count("*") == ctype.count("*"), (name, ctype, rt)
            base = rt.replace("*const ", "").replace("*mut ", "")
            cbase = ctype.replace("const", "").replace("*", "").strip()
            assert width.get(base, base) == cbase, (name, ctype, rt)
            if ctype.count("*") == 1:
                assert rt.startswith("*const ") == ctype.startswith("const "), (name, ctype, rt)
        assert (rret == "") == (ret == "void"), name
    # the safe crate only calls functions that exist
    import re
    safe = open(os.path.join(ROOT, "rust", "lz-fear-b200", "src", "lib.rs")).read()
    for fn in set(re.findall(r"sys::(lzf_\w+)\(", safe)):
        assert fn in rust, fn
    for const in set(re.findall(r"sys::(LZF_\w+)", safe)):
        assert ("pub const %s:" % const) in open(lib_rs).read(), const


def test_rust_struct_layouts_match_the_header():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_sys", os.path.join(ROOT, "rust", "gen_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    from lz_fear_b200 import _native
    # field names and order of the ctypes structures (what the tests exercise) = the header's = the generated Rust
    cs = gen.c_structs()
    assert [f[1] for f in cs["lzf_settings"]] == [f[0] for f in _native.Settings._fields_]
    assert [f[1] for f in cs["lzf_frame_info"]] == [f[0] for f in _native.FrameInfo._fields_]
    assert [f[1] for f in cs["lzf_xxh32_state"]] == [f[0] for f in _native.Xxh32State._fields_]


def test_feature_patch_applies_to_the_reference_snapshot():
    """rust/patches/lz-fear-b200-feature.patch is a real unified diff: `patch --dry-run` accepts every hunk against the
    reference snapshot (skipped where the snapshot is absent, e.g. on the GPU box; a dry run writes nothing)."""
    import shutil
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src", "framed")) or shutil.which("patch") is None:
        pytest.skip("reference snapshot or patch(1) not available")
    p = subprocess.run(["patch", "--dry-run", "-p1", "-d", ref, "-i", os.path.join(ROOT, "rust", "patches", "lz-fear-b200-feature.patch")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("checking file") == 3


def test_rust_sources_are_lexically_balanced():
    """No Rust toolchain here: at least brackets, braces and parentheses of the hand-written crate balance outside
    strings, chars, lifetimes and comments (catches truncated edits; it is not a compile check)."""
    for rel in ("rust/lz-fear-b200/src/lib.rs", "rust/lz-fear-b200-sys/src/lib.rs", "rust/lz-fear-b200-sys/build.rs"):
        src = open(os.path.join(ROOT, rel)).read()
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
        src = re.sub(r"'(?:\\.|[^'\\])'", "' '", src)           # char literals; lifetimes ('a) have no closing quote
        stack, pairs = [], {")": "(", "]": "[", "}": "{"}
        for i, ch in enumerate(src):
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack.pop() == pairs[ch], (rel, src[max(0, i - 80):i + 20])
        assert not stack, (rel, stack[-3:])

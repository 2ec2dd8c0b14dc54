"""Regenerates rust/lz-fear-b200-sys/src/lib.rs from include/lzfear_b200.h (every prototype, every struct, every
constant), so the Rust declarations cannot drift from the C header.  tests/test_abi.py runs the same parser over both
files and compares them symbol by symbol.

    python rust/gen_sys.py            # rewrites src/lib.rs
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lzfear_b200.h")
OUT = os.path.join(ROOT, "rust", "lz-fear-b200-sys", "src", "lib.rs")

SCALARS = {"int": "c_int", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "uint8_t": "u8",
           "void": "c_void", "char": "c_char"}
OPAQUE = ("lzf_ctx", "lzf_table")
STRUCTS = ("lzf_settings", "lzf_frame_info", "lzf_xxh32_state")


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def c_prototypes(text=None):
    """-> [(name, return C type, [(C type, param name)])] in header order."""
    text = strip_comments(text if text is not None else open(HEADER).read())
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    text = re.sub(r"enum\s*\{.*?\}\s*;", " ", text, flags=re.S)
    out = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(lzf_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                mm = re.match(r"(.*?)(\w+)$", p)
                plist.append((mm.group(1).strip(), mm.group(2)))
        out.append((name, ret, plist))
    return out


def rust_type(ctype):
    t = " ".join(ctype.replace("*", " * ").split())
    stars = t.count("*")
    const = t.startswith("const ")
    base = t.replace("const ", "").replace("*", "").strip()
    r = SCALARS.get(base, base)
    for i in range(stars):
        # only the innermost pointer carries the C `const`
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


def c_structs():
    text = strip_comments(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for f in m.group(1).split(";"):
            f = " ".join(f.split())
            if not f:
                continue
            mm = re.match(r"(.*?)(\w+)(\[(\d+)\])?$", f)
            fields.append((mm.group(1).strip(), mm.group(2), int(mm.group(4)) if mm.group(4) else None))
        out[m.group(2)] = fields
    return out


def c_constants():
    text = strip_comments(open(HEADER).read())
    consts = []
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = [x.strip() for x in item.split("=")]
                nxt = int(v, 0)
            else:
                k = item
            consts.append((k, nxt, "i32"))
            nxt += 1
    for m in re.finditer(r"#define\s+(LZF_\w+)\s+(0x[0-9A-Fa-f]+u?|\d+u?)\s", text):
        v = m.group(2).rstrip("u")
        consts.append((m.group(1), int(v, 0), "u32"))
    return consts


def generate():
    L = ["//! Raw FFI declarations of `liblzfear_b200.so` (include/lzfear_b200.h): the B200 LZ4 block codec that stands in for",
         "//! lz-fear's `raw::compress2` / `raw::decompress_raw` and the per-block loops of its framed layer.",
         "//!",
         "//! GENERATED from the C header by rust/gen_sys.py — edit the header, not this file.  Everything here is `unsafe`;",
         "//! the safe surface lives in the sibling crate `lz-fear-b200`.",
         "#![allow(non_camel_case_types)]",
         "use std::os::raw::{c_char, c_int, c_void};", ""]
    for k, v, t in c_constants():
        L.append("pub const %s: %s = %s;" % (k, t, ("0x%X" % v) if v > 65535 else str(v)))
    L.append("")
    for o in OPAQUE:
        L += ["#[repr(C)]", "pub struct %s { _private: [u8; 0] }" % o]
    L.append("")
    for name, fields in c_structs().items():
        L += ["#[repr(C)]", "#[derive(Clone, Copy)]", "pub struct %s {" % name]
        for ct, fn, arr in fields:
            rt = rust_type(ct)
            L.append("    pub %s: %s," % (fn, "[%s; %d]" % (rt, arr) if arr else rt))
        L += ["}", ""]
    L += ['#[link(name = "lzfear_b200")]', 'extern "C" {']
    for name, ret, params in c_prototypes():
        ps = ", ".join("%s: %s" % ({"in": "input", "type": "kind"}.get(pn, pn), rust_type(ct)) for ct, pn in params)
        r = "" if ret == "void" else " -> " + rust_type(ret)
        L.append("    pub fn %s(%s)%s;" % (name, ps, r))
    L += ["}", ""]
    return "\n".join(L)


if __name__ == "__main__":
    open(OUT, "w").write(generate())
    print("wrote", OUT, "(%d functions)" % len(c_prototypes()))

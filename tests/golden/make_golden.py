"""Regenerates tests/golden/* from the read-only reference checkout (default /root/reference).

The reference is Rust and cannot run here, so nothing is *computed* by it: this script only lifts
the byte vectors, strings and corpora its own tests hold for the hot path into language-neutral
files, so that the test-suite does not need /root/reference at run time (it does not exist on the
GPU box).

  reference_vectors.json   decode KATs (src/raw/decompress.rs:153-175), roundtrip strings
                           (src/lib.rs:43-95), big_compression generator parameters (src/lib.rs:97-106)
  issue15_input.bin.xz     the 81 248-byte input of tests/issue-15.rs:5
  corpus.tar.xz            fuzz/corpus/{decode,interop_decode,roundtrip_fuzz}/*
"""
import json
import lzma
import os
import re
import subprocess
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def rust_strings(src, fn_names):
    """inverse("...") string literals of the named #[test] functions in src/lib.rs."""
    out = {}
    for fn in fn_names:
        m = re.search(r"fn %s\(\) \{(.*?)\n    \}" % fn, src, re.S)
        body = m.group(1)
        lits = re.findall(r'inverse\("((?:[^"\\]|\\.)*)"\)', body)
        lits += re.findall(r'let s = "((?:[^"\\]|\\.)*)";', body)
        out[fn] = [bytes(s, "utf-8").decode("unicode_escape") for s in lits]
    return out


def main():
    lib = open(os.path.join(REF, "src/lib.rs")).read()
    strings = rust_strings(lib, ["shakespear", "save_the_pandas", "not_compressible", "short", "empty_string",
                                 "nulls", "compression_works"])
    vectors = {
        "source": "main--/rust-lz-fear: src/raw/decompress.rs:153-175, src/lib.rs:43-106",
        # (input bytes, expected output or null when the reference asserts is_err())
        "decode_kats": [
            {"name": "aaaaaa_no_dup", "input": [0x11, ord("a"), 1, 0], "output": list(b"aaaaaa")},
            {"name": "multiple_repeated_blocks",
             "input": [0x11, ord("a"), 1, 0, 0x22, ord("b"), ord("c"), 2, 0], "output": list(b"aaaaaabcbcbcbc")},
            {"name": "all_literal", "input": [0x30, ord("a"), ord("4"), ord("9")], "output": list(b"a49")},
            {"name": "offset_oob_1", "input": [0x10, ord("a"), 2, 0], "output": None},
            {"name": "offset_oob_2", "input": [0x40, ord("a"), 1, 0], "output": None},
        ],
        "roundtrip_strings": strings,
        "big_compression": {"n": 80000000, "formula": "((n as u8) * 0xA + 33) ^ 0xA2 (wrapping)"},
    }
    with open(os.path.join(OUT, "reference_vectors.json"), "w") as f:
        json.dump(vectors, f, indent=1)

    issue = open(os.path.join(REF, "tests/issue-15.rs")).read()
    arr = re.search(r"let input = \[(.*?)\];", issue, re.S).group(1)
    data = bytes(int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", arr))
    assert len(data) == 81248, len(data)
    with open(os.path.join(OUT, "issue15_input.bin.xz"), "wb") as f:
        f.write(lzma.compress(data, preset=9))

    subprocess.check_call(["tar", "--sort=name", "--mtime=2020-01-01", "--owner=0", "--group=0", "--numeric-owner",
                           "-cJf", os.path.join(OUT, "corpus.tar.xz"), "-C", os.path.join(REF, "fuzz/corpus"),
                           "decode", "interop_decode", "roundtrip_fuzz"])
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()

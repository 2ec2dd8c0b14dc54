#!/usr/bin/env python
"""dolz4 <in> <out> — the reference's examples/dolz4.rs:14-17 on the B200 codec:
CompressionSettings::default().compress_with_size(file_in, file_out)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lz_fear_b200 as lz  # noqa: E402


def main():
    if len(sys.argv) != 3:
        raise SystemExit("usage: dolz4.py <in> <out>")
    with open(sys.argv[1], "rb") as fin, open(sys.argv[2], "wb") as fout:
        lz.CompressionSettings.default().compress_with_size(fin, fout)


if __name__ == "__main__":
    main()

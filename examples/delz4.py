#!/usr/bin/env python
"""delz4 <in> <out> — the reference's examples/delz4.rs:13-20 on the B200 codec:
LZ4FrameReader::new(file)?.into_read(), then a fill_buf / consume loop (one block per fill_buf)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lz_fear_b200 as lz  # noqa: E402


def main():
    if len(sys.argv) != 3:
        raise SystemExit("usage: delz4.py <in> <out>")
    with open(sys.argv[1], "rb") as fin, open(sys.argv[2], "wb") as fout:
        reader = lz.LZ4FrameReader(fin).into_read()
        while True:
            buf = reader.fill_buf()
            if not buf:
                break
            fout.write(buf)
            reader.consume(len(buf))


if __name__ == "__main__":
    main()

"""Runs a slice of the parity checks against the AddressSanitizer build of the SIMT-emulated product
sources (out-of-bounds reads/writes of the kernels and of the C-ABI layer show up as ASan reports).
Started by tests/test_simt_asan.py with LD_PRELOAD=libasan."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
import parity  # noqa: E402
from lz_fear_b200 import _native as N  # noqa: E402


class B:
    pass


def main():
    N.load_library(os.path.join(ROOT, "tests", "simt", "libsimt_lzfear_asan.so"))
    b = B()
    b.ctx = N.Context(0)
    inputs = [d for d in parity.sample_inputs() if len(d) <= 70000]
    parity.check_raw_compress(b, oracle, inputs[::2])
    blocks = [(oracle.compress_block(d)[1], len(d)) for d in inputs]
    parity.check_raw_decompress(b, oracle, blocks[::2])
    bad = [(parity.mutate(c, i, k=2), None) for i, (c, _n) in enumerate(blocks) if len(c) > 8]
    parity.check_raw_decompress(b, oracle, bad)
    parity.check_frames(b, oracle, [b"", bytes(65536), inputs[6], inputs[5] * 3], parity.FRAME_SETTINGS[:3])
    frames = []
    for kw in parity.FRAME_SETTINGS[1:3]:
        rc, frame = oracle.frame_compress(inputs[6] + inputs[5], **kw)
        frames += [parity.mutate(frame, 7 * k, k=2) for k in range(10)]
    parity.check_frame_decode_errors(b, oracle, frames)
    # a slice of the seeded structural fuzz (the whole campaign ran clean under this build once)
    parity.check_fuzz_blocks(b, oracle, seed=301, count=16, max_len=40000)
    parity.check_fuzz_frame_batches(b, oracle, seed=303, count=4, max_len=30000)
    parity.check_table_limit_zone(b, oracle, seed=9, count=16)
    print("ASAN-RUN-OK")


if __name__ == "__main__":
    main()

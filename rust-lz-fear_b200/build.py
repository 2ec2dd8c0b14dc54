"""Builds liblzfear_b200.so (the C-ABI shared library of include/lzfear_b200.h) IN-TREE with nvcc
for sm_100a.  nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.

    python rust-lz-fear_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblzfear_b200.so")
SOURCES = ["lzf_api.cu", "lzf_compress.cu", "lzf_decompress.cu", "lzf_frame.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "--shared", "-cudart", "shared",
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _deps():
    out = [os.path.join(HERE, "..", "include", "lzfear_b200.h"), os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh")):
            out.append(os.path.join(CSRC, f))
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (e.g. defines=["LZF_DEC_MINCTAS=4"], out="build/variant.so")."""
    target = out or LIB
    if not force and not defines and out is None and up_to_date():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(os.path.abspath(target)), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", os.path.abspath(target)] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))

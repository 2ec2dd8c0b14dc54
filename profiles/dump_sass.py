"""Writes profiles/r02_sass_{encode,decode}.txt: the SASS of the two block kernels of the shipped library
(`cuobjdump -xelf` + `nvdisasm -c`), with an opcode histogram on top (UBLKCP / SYNCS = TMA bulk copies and mbarriers,
MATCH / VOTE / SHFL = the warp collectives of the parse)."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rust-lz-fear_b200", "liblzfear_b200.so")


def extract(sass, substr, out, title):
    on, buf, ops = False, [], {}
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            on = substr in ln
        if on:
            buf.append(ln)
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                op = m.group(1).split(".")[0]
                ops[op] = ops.get(op, 0) + 1
    hdr = ["# %s" % title,
           "# nvdisasm -c of the shipped rust-lz-fear_b200/liblzfear_b200.so (sm_100a); %d SASS instructions" % sum(ops.values()),
           "# opcode histogram: " + ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:40]), ""]
    open(out, "w").write("\n".join(hdr + buf) + "\n")
    print(out, sum(ops.values()))


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True, check=True)
    for unit, substr, name, title in (
            ("lzf_compress", "encode_blocks_kernelILi2ELb0ELi28", "encode", "encode_blocks_kernel<2,false,28> (packed 17-bit tables, 28-warp CTA): the config-3 instantiation"),
            ("lzf_decompress", "decode_blocks_kernel", "decode", "decode_blocks_kernel")):
        cub = [f for f in os.listdir(tmp) if f.startswith(unit + ".")][0]
        sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        extract(sass, substr, os.path.join(ROOT, "profiles", "r02_sass_%s.txt" % name), title)


if __name__ == "__main__":
    main()

"""Per-source-line view of an `ncu --set full --import-source on` capture: joins the SASS rows of the
report's source page with the line table of the shipped cubin (nvdisasm -g), by instruction offset.

    python profiles/line_profile.py gpurun_out/prof_encode_v5.ncu-rep encode_blocks_kernelILi2 lzf_compress [top]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_substr, cubin_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("LZF_PROFILE_LIB") or os.path.join(ROOT, "rust-lz-fear_b200", "liblzfear_b200.so")], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith(cubin_substr + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    out, on, cur = {}, False, ("?", 0)
    for ln in dis:
        if ln.startswith(".text."):
            on = kernel_substr in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    rep, kern, cub = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    lt = line_table(kern, cub)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ismp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    base = int(rows[2][ia], 16)
    per = collections.defaultdict(lambda: [0, 0, 0])
    tot_ex = tot_s = 0
    for r in rows[2:]:
        if len(r) <= iex or not r[iex].isdigit():
            continue
        off = int(r[ia], 16) - base
        key = lt.get(off, ("?", 0))
        per[key][0] += int(r[iex]); per[key][1] += int(r[ismp]); per[key][2] += 1
        tot_ex += int(r[iex]); tot_s += int(r[ismp])
    print("total warp instructions %d, samples %d" % (tot_ex, tot_s))
    print("%-22s %8s %8s %6s" % ("file:line", "instr%", "stall%", "sass"))
    for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %7.2f%% %7.2f%% %6d" % ("%s:%d" % key, 100.0 * v[0] / tot_ex, 100.0 * v[1] / max(tot_s, 1), v[2]))


if __name__ == "__main__":
    main()

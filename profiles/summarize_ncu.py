"""Turns an .ncu-rep (from `ncu --set full`) into the short text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_decode_v3.ncu-rep profiles/r01_decode_v3.txt "note"
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + rep.split("/")[-1] + (" — " + note if note else ""), ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append("kernel: " + name)
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                lines.append("  %-82s %-14s %s" % (h, u, v))
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        data = [(int(r[ismp]), int(r[iex]), r[isrc].strip()) for r in rows[2:] if len(r) > ismp and r[ismp].isdigit()]
        tot = sum(d[0] for d in data) or 1
        lines.append("top stall-sample instructions (share of samples, warp-level executions, SASS):")
        for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:20]):
            lines.append("  #%-5d %5.1f%%  %12d  %s" % (i, 100.0 * data[i][0] / tot, data[i][1], data[i][2][:90]))
        ops = {}
        for d in data:
            op = d[2].split()[0] if not d[2].startswith("@") else d[2].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + d[1]
        totex = sum(ops.values()) or 1
        lines.append("")
        lines.append("executed warp instructions by opcode: " +
                     ", ".join("%s %.1f%%" % (k, 100.0 * v / totex) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

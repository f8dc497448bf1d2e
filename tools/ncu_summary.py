#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) of one kernel: duration, DRAM traffic, pipe utilisation,
stall reasons and where (by opcode) the warp-stall samples fall.  Usage: tools/ncu_summary.py rep [out.txt]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    lines = []
    rows = page(rep, "raw")
    hdr, units, val = rows[0], rows[1], rows[-1]
    m = {h: (u, v) for h, u, v in zip(hdr, units, val)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
    for w in want:
        if w in m:
            lines.append("%-80s %s %s" % (w, m[w][1], m[w][0]))
    st = sorted(((float(v[1]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)
    lines.append("-- stall reasons (warps stalled per issue-active cycle)")
    for v, k in st[:8]:
        lines.append("   %6.3f %s" % (v, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    src = page(rep, "source")
    h2 = src[1]
    isrc, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    data = src[2:]
    tot = sum(int(r[isamp]) for r in data) or 1
    c, ex = Counter(), Counter()
    for r in data:
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        c[op] += int(r[isamp]); ex[op] += int(r[iex])
    lines.append("-- warp-stall samples by opcode (%d SASS instructions, %d samples)" % (len(data), tot))
    for op, v in c.most_common(12):
        lines.append("   %-8s %5.1f%%   executed %d" % (op, 100.0 * v / tot, ex[op]))
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()

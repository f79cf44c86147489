#!/usr/bin/env python3
"""Summarise an ncu report (run here, no GPU needed): key raw metrics + instructions per source line.

usage: python tools/ncu_summary.py gpurun_out/foo.ncu-rep [--top 30] [--kernel REGEX] > profiles/foo_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_branch.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def run(args):
    # --kernel <regex> (optional) selects one kernel of a multi-kernel report
    if "--kernel" in sys.argv:
        args = args + ["-k", "regex:" + sys.argv[sys.argv.index("--kernel") + 1]]
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"# {rep}\n## raw metrics")
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:120])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {vals[i]} {units[i]}")
    print("## warp stall reasons (warps per issue-active cycle)")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            print(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {float(vals[i]):.3f}")
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, agg = None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[7]), int(r[4])))
            except ValueError:
                pass
    tot = sum(a[3] for a in agg) or 1
    smp = sum(a[4] for a in agg) or 1
    print(f"## instructions executed per source line (total {tot / 1e9:.2f} G warp-instructions, {smp} stall samples)")
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print(f"{a[0]:14s}:{a[1]:<4d} {a[3] / 1e9:7.2f}G {100 * a[3] / tot:5.1f}%  samples {100 * a[4] / smp:5.1f}% | {a[2]}")
    byfile = {}
    for a in agg:
        byfile[a[0]] = byfile.get(a[0], 0) + a[3]
    print("## by file:", {k: f"{v / 1e9:.2f}G" for k, v in byfile.items()})


if __name__ == "__main__":
    main()

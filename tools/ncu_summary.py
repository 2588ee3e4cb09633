"""Summarise an .ncu-rep (read here, no GPU needed) into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass]
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    gi = hdr.index("Grid Size")
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"\n## {r[ki][:90]}  grid {r[gi]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:85s} {r[i]:>16s} {rows[1][i]}")
    if "--sass" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        hdr = rows[1]
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        ops, samp = collections.Counter(), collections.Counter()
        for r in rows[2:]:
            if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
                break
            t = r[ia].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += int(r[ie])
            samp[op] += int(r[isamp])
        tot, tots = sum(ops.values()), max(sum(samp.values()), 1)
        print(f"\n## SASS instruction mix of the first kernel ({tot} warp instructions, {tots} samples)")
        for op, c in ops.most_common(16):
            print(f"{op:10s} {100 * c / tot:5.1f}% of instructions   {100 * samp[op] / tots:5.1f}% of stall samples")


if __name__ == "__main__":
    main()

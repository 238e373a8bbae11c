#!/usr/bin/env python
"""Digest an `ncu --set full` report into the few counters the roofline discussion needs (markdown).

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [title] > profiles/x_digest.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("Source: `%s` (ncu --set full --clock-control none; one row per captured launch)\n" % rep)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print("## %s  (launch id %s)\n" % (name.split("(")[0].replace("<unnamed>::", ""), r[col["ID"]]))
        print("| counter | value | unit |")
        print("|---|---:|---|")
        for k in KEYS:
            if k in col:
                print("| %s | %s | %s |" % (k, r[col[k]], units[col[k]]))
        print()


if __name__ == "__main__":
    main()

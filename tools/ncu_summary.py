#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into a small CSV for profiles/:
one row per profiled launch with duration, DRAM bytes, hit rates, occupancy, registers."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{n} [{units[i]}]" if units[i] else n for n, i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split("(")[0] if n == "Kernel Name" else r[i] for n, i in idx])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

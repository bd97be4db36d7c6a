#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): headline metrics + SASS hot segments. Usage: ncu_summary.py file.ncu-rep [--src]"""
import collections, csv, io, subprocess, sys

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]
    return {n: (v[i], u[i]) for i, n in enumerate(h)}

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]

def main():
    path = sys.argv[1]
    d = raw(path)
    for k in KEYS:
        if k in d:
            print(f"{k}: {d[k][0]} {d[k][1]}")
    if "--src" in sys.argv:
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))[2:]
        tot = sum(int(r[5]) for r in rows); ts = sum(int(r[4]) for r in rows)
        print("total warp-instructions", tot, "samples", ts)
        seg = []
        for i, r in enumerate(rows):
            c, s = int(r[5]), int(r[4])
            if seg and seg[-1][2] == c: seg[-1][1] = i; seg[-1][3] += c; seg[-1][4] += s
            else: seg.append([i, i, c, c, s])
        for a, b, c, t, s in seg:
            if t / max(tot, 1) > 0.015 or s / max(ts, 1) > 0.03:
                oc = collections.Counter(rows[k][1].split()[0] if not rows[k][1].split()[0].startswith('@') else rows[k][1].split()[1] for k in range(a, b + 1)).most_common(6)
                print("%4d-%4d n=%3d exec=%9d inst%%=%5.1f samp%%=%5.1f %s" % (a, b, b - a + 1, c, 100 * t / tot, 100 * s / ts, oc))
        for r in sorted(rows, key=lambda r: -int(r[4]))[:8]:
            print("  hot:", r[4], r[1].strip()[:80])

if __name__ == "__main__":
    main()

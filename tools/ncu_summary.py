#!/usr/bin/env python
"""Summarises ncu outputs into profiles/ (run here, no GPU needed).
  tools/ncu_summary.py launches gpurun_out/launches_<tag>.csv   -> per-kernel time shares
  tools/ncu_summary.py full gpurun_out/sweep_<tag>.ncu-rep        -> key metrics of the capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_drain_per_warp_active.pct", "launch__func_cache_config", "launch__shared_mem_per_block_dynamic"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, R = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    d = collections.defaultdict(list)
    for r in R:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("# kernel launches of `python bench.py --steps 20 --warmup 3 --no-cpu-baseline` under")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised; compare SHARES)")
    print("%-70s %6s %14s %8s %12s" % ("kernel", "n", "total_ns", "share", "avg_ns"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%-70s %6d %14.0f %8.4f %12.0f" % (k[:70], len(v), sum(v), sum(v) / tot, sum(v) / len(v)))


def full(path):
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        print("# kernel:", V[H.index("Kernel Name")], " grid", V[H.index("Grid Size")], " block", V[H.index("Block Size")])
        for i, h in enumerate(H):
            if h in KEYS:
                print("%-80s %-16s %s" % (h, U[i], V[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep                > profiles/rNN_kernel.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ki].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); cold-cache serialised launches: compare shares")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} n={v[0]:5d} total_ms={v[1] / 1e6:11.3f} avg_us={v[1] / v[0] / 1e3:10.2f} share={v[1] / tot:6.1%}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none  ({path})")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:160])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:80s} {r[i]:>20s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

#!/usr/bin/env python
"""Condense an .ncu-rep into the short text summaries kept under profiles/ (reads the report here, no GPU needed).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep "comment line" ... > profiles/rNN_....txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    for c in sys.argv[2:]:
        print("# " + c)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print(f"{'Kernel Name':<88} {d.get('Kernel Name', '')}")
        for k in KEEP:
            if k in d:
                print(f"{k:<88} {d[k]} {u[k]}")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v),
                        reverse=True)
        for v, k in stalls[:8]:
            print(f"{k:<88} {v:.6f} inst")
        print()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""
Summarise an Nsight Compute report (.ncu-rep, brought back from the GPU box in gpurun_out/)
into the few numbers DESIGN.md / bench.py quote:  duration, DRAM bytes, issue utilisation,
pipe utilisation, occupancy limiters, top stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_<kernel>.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"kernel: {name}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]} {units[i]}")
        for i, h in enumerate(hdr):
            is_stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
            is_pipe = h.startswith("sm__inst_executed_pipe") and h.endswith(".avg.pct_of_peak_sustained_active")
            if is_stall or is_pipe:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.3:
                    print(f"  {h:72s} {r[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])

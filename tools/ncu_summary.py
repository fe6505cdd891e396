#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i`, no GPU needed) into a small text summary for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        print(f"--- launch {n}: {name[:110]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:70s} {r[i]} {units[i]}")
    stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_not_issued.ratio") is False]
    if rows[2:] and stall:
        r = rows[2]
        vals = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall if r[hdr.index(h)].replace(".", "").isdigit()), reverse=True)[:8]
        print("--- top warp-stall reasons (launch 0, warps per issue slot):")
        for v, h in vals:
            print(f"    {h:90s} {v:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])

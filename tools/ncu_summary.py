#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / the roofline objects cite.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import re
import subprocess
import sys

PATTERNS = [
    r"^gpu__time_duration\.sum$", r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|shared_mem_per_block_\w+)$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$", r"^smsp__thread_inst_executed_per_inst_executed\.ratio$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|xu|lsu|adu|cbu|uniform|tc|tensor\w*|tmem|tma)\.avg\.pct_of_peak_sustained_active$",
    r"^sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$", r"^sm__pipe_tc_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$",
    r"^l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit\.sum$", r"^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$",
    r"^smsp__warps_eligible\.avg\.per_cycle_active$", r"^sm__cycles_active\.avg$", r"^smsp__inst_executed_op_(tma_ld|ldgsts)\.sum$",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u in zip(hdr, units):
            if any(re.search(p, h) for p in PATTERNS):
                v = d[h]
                if "stalled" in h and v not in ("", "0") and float(v) < 0.05:
                    continue
                print(f"  {h:90s} {v:>18s} {u}")
        print()


if __name__ == "__main__":
    main()

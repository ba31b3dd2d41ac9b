#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (read on the CPU box): python tools/ncu_summary.py <file.ncu-rep> [regex ...]"""
import csv
import re
import subprocess
import sys

WANT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^launch__registers_per_thread$", r"^launch__occupancy_limit_(shared_mem|registers|warps)$",
        r"^launch__shared_mem_per_block_dynamic$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$", r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st)\.sum$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared_op_(ld|st)\.sum$",
        r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", r"^smsp__average_warp_latency_issue_stalled_(long_scoreboard|short_scoreboard|barrier|membar|wait|lg_throttle|math_pipe_throttle|mio_throttle|sleeping|no_instruction|not_selected|selected|dispatch_stall|branch_resolving|drain|imc_miss|tex_throttle)\.ratio$"]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(p) for p in WANT + extra]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("## %s  grid %s block %s" % (r[ki][:100], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "", r[hdr.index("Block Size")] if "Block Size" in hdr else ""))
        for h, u, v in zip(hdr, units, r):
            if any(p.search(h) for p in pats) and v not in ("", "0"):
                print("  %-90s %12s %s" % (h, v, u))


if __name__ == "__main__":
    main()

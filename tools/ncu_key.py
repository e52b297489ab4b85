#!/usr/bin/env python
"""Print the handful of metrics that decide what bounds a kernel from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active" ]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", rep, r[hdr.index("Kernel Name")][:70])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-70s %s %s" % (w, r[i], units[i]))
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
        stalls.sort(reverse=True)
        for v, h in stalls[:6]:
            print("  stall %-62s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))

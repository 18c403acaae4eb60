#!/usr/bin/env python3
"""ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) -> the small tracked text summary used under profiles/.
usage: python profiles/ncu_text.py <raw.csv> <out.txt> "<command line that was profiled>" """
import csv
import sys

KEEP = ["gpu__time_duration.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on   %s\n# (one block per captured launch; times under the profiler are cold-cache and serialised)\n" % sys.argv[3])
    for r in rows[2:]:
        f.write("\n== %s  grid=%s block=%s\n" % (r[hdr.index("Kernel Name")].split("(")[0], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?",
                                               r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    if float(r[i].replace(",", "")) != 0:
                        f.write("   %-84s %16s %s\n" % (h, r[i], units[i]))
                except ValueError:
                    pass

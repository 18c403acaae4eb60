#!/usr/bin/env python3
"""Turn gpurun_out artefacts into the small, tracked summaries under profiles/.
usage: python profiles/summarize.py <tag>      (reads gpurun_out/*_<tag>.*, writes profiles/<tag>_*.txt|json)"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

TAG = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def launches():
    path = os.path.join(G, "launches_%s.csv" % TAG)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > vi:
            try:
                agg[r[ki].split("(")[0][:70]].append(float(r[vi].replace(",", "")) / 1e3)
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, "%s_launches.txt" % TAG), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 4 --warmup 3 --skip-cpu\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write("%-72s %5s %12s %12s %7s\n" % ("kernel", "n", "mean_us", "total_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-72s %5d %12.1f %12.1f %6.1f%%\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))


def full(name):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (name, TAG))
    if not os.path.exists(rep):
        return {}
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    with open(os.path.join(P, "%s_ncu_%s.txt" % (TAG, name)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (one row per captured launch)\n")
        for r in rows[2:]:
            kn = r[hdr.index("Kernel Name")].split("(")[0]
            f.write("== %s  grid=%s block=%s\n" % (kn, r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?", r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
            for i, h in enumerate(hdr):
                if h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        if float(r[i].replace(",", "")) != 0:
                            f.write("   %-84s %16s %s\n" % (h, r[i], units[i]))
                    except ValueError:
                        pass
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = rd * mult[units[hdr.index("dram__bytes_read.sum")]] + wr * mult[units[hdr.index("dram__bytes_write.sum")]]
                res.setdefault(kn, []).append(tr)
            except Exception:
                pass
    return res


launches()
traffic = {}
for nm in ("bpr", "score", "als"):
    for k, v in full(nm).items():
        traffic[k] = sum(v) / len(v)
tj = os.path.join(P, "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
grad = next((v for k, v in traffic.items() if "bpr_grad" in k), None)
app = next((v for k, v in traffic.items() if "bpr_apply" in k), None)
sc = next((v for k, v in traffic.items() if "score_topk" in k or "score_filter" in k), None)
if grad is not None and app is not None:
    old["bpr_step"] = grad + app
if sc is not None:
    old["score_topk"] = sc
old["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, tag %s" % TAG
json.dump(old, open(tj, "w"), indent=1)
b = os.path.join(G, "bench_%s.json" % TAG)
if os.path.exists(b):
    open(os.path.join(P, "%s_bench.json" % TAG), "w").write(open(b).read())
print(open(tj).read())

#!/bin/bash
# round 2, GPU call F (1 GPU): VBPR with hot items + GEMM variants; per-kernel launch list of a VBPR step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_vbpr.py tests/test_gpu_bpr.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python profiles/probe_gemm3.py 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_vbpr_r02f.csv python profiles/run_vbpr.py 20 2 > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows = list(csv.reader(open("gpurun_out/launches_vbpr_r02f.csv")))
st = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
h = rows[st]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = defaultdict(list)
for r in rows[st + 1:]:
    if len(r) > vi:
        try: agg[r[ki].split("(")[0][:60]].append(float(r[vi].replace(",", "")) / 1e3)
        except ValueError: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print("%-62s n=%3d mean %9.1f us" % (k, len(v), sum(v) / len(v)))
PY

#!/bin/bash
# ALS section of the bench alone (GPU idle before it) and inside the full line (after >= 1 s of tensor-pipe work), with its clocks
mkdir -p gpurun_out
timeout 600 python bench.py --skip-score --skip-sweep --skip-cpu > gpurun_out/bench_final2_alsonly.json 2> /dev/null
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -2 gpurun_out/bench_final2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_final2_alsonly.json", "gpurun_out/bench_final2.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1]); a = d["als"]
    print(f, round(a["user_step_ms"], 1), round(a["item_step_ms"], 1), a.get("clocks"), round(d["value"] / 1e9, 3), d["roofline"]["frac"])
PY

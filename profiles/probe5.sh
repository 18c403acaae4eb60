#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_gpu_topk.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import torch, bench
peaks, _ = bench.measured_peaks()
for r in bench.score_sweep(torch.device('cuda'), 18944, 1 << 20, 30, peaks): print(r)
PY

#!/bin/bash
# round 2, GPU call E (1 GPU): tcgen05 3xTF32 content GEMMs of VBPR -- tests, then timing at C3
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_vbpr.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_r02e.log; cat gpurun_out/pytest_r02e.log
timeout 600 python -m pytest tests/test_gpu_baseline.py -m gpu -q -x -k c3 2>&1 | tail -8
timeout 300 python - <<'PY'
import sys, os, json
sys.path[:0] = ["top-k-rec_b200", "."]
import torch, bench, topkrec
dev = torch.device("cuda", 0)
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
print(json.dumps(bench.vbpr_points(smp, dev, 64.0)))
PY

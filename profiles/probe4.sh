#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_gpu_bpr.py -m gpu -q -x 2>&1 | tail -15
timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import torch, bench, topkrec, ctypes
L = topkrec.lib(); L.tkr_debug_set_count_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_count_mode.restype = None
dev = torch.device('cuda')
for mode in (0, -1):
    L.tkr_debug_set_count_mode(mode)
    for (nu, ni, B) in ((6_000_000, 1_000_000, 1 << 20), (4_000_000, 4_000_000, 1 << 18), (4_000_000, 4_000_000, 1 << 20)):
        print('count_mode', mode, json.dumps(bench.bpr_hbm_streaming(dev, nu, ni, B)), flush=True)
PY

#!/bin/bash
# round 2, GPU call X (1 GPU): K1 tests with the automatic routing between the two multi-step kernels (BPR tests + the BASELINE-shape tests)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_bpr.py tests/test_gpu_baseline.py -q -x 2>&1 | tail -8 > gpurun_out/pytest_r02x.log; cat gpurun_out/pytest_r02x.log

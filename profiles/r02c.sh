#!/bin/bash
# round 2, GPU call C (1 GPU): persistent B=256 kernel (tests + timing), fixed C2 replay test, K3 shard-size probe
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_bpr.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_r02c.log; cat gpurun_out/pytest_r02c.log
timeout 600 python -m pytest tests/test_gpu_baseline.py -m gpu -q -x -k "c2" 2>&1 | tail -8 > gpurun_out/pytest_r02c_c2.log; cat gpurun_out/pytest_r02c_c2.log
timeout 300 python profiles/probe_b256.py 2>&1 | tail -50
timeout 300 python profiles/probe_shard.py 2>&1 | tail -20

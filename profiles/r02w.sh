#!/bin/bash
# round 2, GPU call W (1 GPU): the dataflow multi-step kernel (no grid barriers) -- parity tests, then us per step against the
# cluster kernel and the two-launch route
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_bpr.py -q -x -k "dataflow" 2>&1 | tail -15 > gpurun_out/pytest_r02w.log; cat gpurun_out/pytest_r02w.log
timeout 600 python profiles/probe_b256.py 2>&1 | tail -50 > gpurun_out/probe_b256_r02w.txt; grep -v "fused_sampler': True" gpurun_out/probe_b256_r02w.txt | cut -c1-200

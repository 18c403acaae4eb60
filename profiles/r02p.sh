#!/bin/bash
# round 2, GPU call P (1 GPU): blocked ALS factorisation -- parity tests, then the timing probe (legacy loops vs blocked rounds)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_als.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_r02p.log; cat gpurun_out/pytest_r02p.log
for mode in 2 0; do timeout 300 python profiles/als_probe.py 0.125 256 4096 144 $mode 2>&1 | tail -4; done > gpurun_out/als_probe_r02p.txt; cat gpurun_out/als_probe_r02p.txt
for mode in 2 1; do timeout 300 python profiles/als_probe.py 0.125 128 4096 144 $mode 2>&1 | tail -2; timeout 300 python profiles/als_probe.py 0.125 192 4096 144 $mode 2>&1 | tail -2; done > gpurun_out/als_probe_r02p_small.txt; cat gpurun_out/als_probe_r02p_small.txt

#!/bin/bash
# round 2, GPU call P (1 GPU): blocked ALS factorisation -- parity tests, then the timing probe (mode 2 = the loops of round 1, 0 = blocked rounds)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_als.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_r02p.log; cat gpurun_out/pytest_r02p.log
for d in 256 192 128 64; do for mode in 2 0; do echo "d=$d mode=$mode"; timeout 300 python profiles/als_probe.py 0.125 $d 4096 144 $mode 2>&1 | tail -1; done; done > gpurun_out/als_probe_r02p.txt; cat gpurun_out/als_probe_r02p.txt

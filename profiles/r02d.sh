#!/bin/bash
# round 2, GPU call D (1 GPU): persistent kernel after the fence removal, then the new bench.py end to end
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_bpr.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_r02d.log; cat gpurun_out/pytest_r02d.log
timeout 300 python profiles/probe_b256.py 2>&1 | grep -v two_kernel | tail -12
timeout 900 python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; tail -c 6000 gpurun_out/bench_r02d.json; tail -5 gpurun_out/bench_r02d.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400

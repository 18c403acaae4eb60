#!/bin/bash
# round 2, GPU call Y (1 GPU): dataflow kernel with the fused sampler, memcheck on the dataflow tests, bench sweep points
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bpr.py -q -x -k "dataflow" 2>&1 | tail -6 > gpurun_out/pytest_r02y.log; cat gpurun_out/pytest_r02y.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_bpr.py -q -x -k "dataflow and not 70000" 2>&1 | tail -6 > gpurun_out/sanitizer_memcheck_flow_r02y.log; cat gpurun_out/sanitizer_memcheck_flow_r02y.log

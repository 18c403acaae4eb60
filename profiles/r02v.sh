#!/bin/bash
# round 2, GPU call V (1 GPU): compute-sanitizer memcheck over the GPU suite (BASELINE-shape and full-size tests left out: minutes each under the tool)
mkdir -p gpurun_out
timeout 840 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_baseline.py -k "not full_size and not c2 and not 1000_steps and not fold0" 2>&1 | tail -40 > gpurun_out/sanitizer_memcheck_suite_r02v.log; tail -12 gpurun_out/sanitizer_memcheck_suite_r02v.log

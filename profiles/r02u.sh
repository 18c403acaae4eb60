#!/bin/bash
# round 2, GPU call U (1 GPU): compute-sanitizer on the blocked ALS factorisation (memcheck: out-of-bounds / misaligned shared and
# global accesses; racecheck: shared-memory hazards between the pivot warp, the row solvers and the tile warps)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_als.py -q -x -k "both_factorisations and 1- or degenerate or split_rows" 2>&1 | tail -25 > gpurun_out/sanitizer_memcheck_r02u.log; tail -8 gpurun_out/sanitizer_memcheck_r02u.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 400 python -m pytest tests/test_gpu_als.py -q -x -k "both_factorisations and 1-256-48 or both_factorisations and 1-128-100 or both_factorisations and 1-12-16" 2>&1 | tail -600 > gpurun_out/sanitizer_racecheck_r02u.log; tail -25 gpurun_out/sanitizer_racecheck_r02u.log

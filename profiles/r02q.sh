#!/bin/bash
# round 2, GPU call Q (1 GPU): ncu --set full of the blocked ALS kernel (source-level stall sampling), user half-step at d=256
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_segment_kernel -s 5 -c 1 -f -o gpurun_out/prof_als_r02q \
    python profiles/als_probe.py 0.04 256 4096 144 0 > gpurun_out/ncu_als_r02q.log 2>&1
tail -3 gpurun_out/ncu_als_r02q.log

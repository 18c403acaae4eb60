#!/bin/bash
# ALS probe on one box: round-1 loops (mode 2) vs blocked rounds (mode 0), probe shape and bench shape (mean_pos 208)
mkdir -p gpurun_out
for mp in 144 208; do for mode in 2 0; do echo "d=256 mean_pos=$mp mode=$mode"; timeout 300 python profiles/als_probe.py 0.125 256 4096 $mp $mode 2>&1 | tail -1; done; done > gpurun_out/als_probe_r02p2.txt; cat gpurun_out/als_probe_r02p2.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv

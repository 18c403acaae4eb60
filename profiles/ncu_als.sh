#!/bin/bash
# ncu --set full capture of the ALS segment kernel (user side, d=256) + launch list; run under gpurun
set -x
TAG=${1:-r01p}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:als_segment_kernel -s ${SKIP:-5} -c 1 -f -o gpurun_out/prof_als_$TAG \
    python profiles/als_probe.py 0.02 256 4096 208 > gpurun_out/ncu_als_$TAG.log 2>&1
ncu -i gpurun_out/prof_als_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_als_${TAG}_raw.csv 2>/dev/null

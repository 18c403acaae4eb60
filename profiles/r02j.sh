#!/bin/bash
# round 2, GPU call J (1 GPU): ncu --set full captures of the round's new kernels + the filter kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
cap() {   # name, kernel regex, skip, count, command...
  n=$1; k=$2; s=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_$n "$@" > gpurun_out/ncu_$n.log 2>&1
  ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_raw.csv 2>/dev/null
  ls -la gpurun_out/prof_$n.ncu-rep | awk '{print $5, $9}'
}
cap gemm3 gemm_tf32x3 2 2 python profiles/run_vbpr.py 20 2
cap persist bpr_persist 1 1 python profiles/probe_persist.py
cap filter score_filter 2 1 python profiles/run_shard.py 20 3
cap filter_shard score_filter 2 1 python profiles/run_shard.py 17 3
rm -f gpurun_out/prof_persist.ncu-rep gpurun_out/prof_filter_shard.ncu-rep

#!/bin/bash
# TMEM read-bandwidth microbenchmark + filter ceiling probes (one gpurun call)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 profiles/ubench/tmem_ld > gpurun_out/ubench_tmem_ld.txt 2>&1
cat gpurun_out/ubench_tmem_ld.txt | grep "grid=148"
timeout 300 python profiles/dbg_filter.py 128 2>&1 | tee gpurun_out/dbg_filter_d128.txt
timeout 300 python profiles/dbg_filter.py 256 2>&1 | tee gpurun_out/dbg_filter_d256.txt
timeout 300 python profiles/dbg_filter.py 64 2>&1 | tee gpurun_out/dbg_filter_d64.txt

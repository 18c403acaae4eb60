#!/bin/bash
# round 2, GPU call S (1 GPU): the bench line with the round's final kernels (streaming-route register fix, blocked ALS)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err; tail -c 600 gpurun_out/bench_r02s.json; tail -3 gpurun_out/bench_r02s.err

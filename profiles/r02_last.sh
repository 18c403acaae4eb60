#!/bin/bash
# last GPU call of the round (1 GPU): whole GPU suite + smoke + bench line after the filter kernel's third tensor map
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_last.log; cat gpurun_out/pytest_last.log
timeout 900 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -c 300 gpurun_out/bench_last.json; tail -2 gpurun_out/bench_last.err

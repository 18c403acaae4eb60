#!/bin/bash
# one GPU: segment tests, then per-segment times of a sweep cut into 8 / 4 / 2 segments (what a ring slot costs without any peer traffic)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_topk.py -q -x -k "segments" 2>&1 | tail -5
for G in 8 4 2; do timeout 300 python profiles/probe_segments.py $G 20; done > gpurun_out/probe_segments.json; cat gpurun_out/probe_segments.json

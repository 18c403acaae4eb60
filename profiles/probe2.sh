#!/bin/bash
# correctness of the tensor-core path vs the oracle + filter role counters (one gpurun call)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python profiles/dbg_tc.py 2>&1 | tee gpurun_out/dbg_tc.txt
timeout 200 python profiles/dbg_filter.py 128 2>&1 | cat | tee gpurun_out/dbg_filter_d128.txt

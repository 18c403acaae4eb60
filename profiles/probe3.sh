#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python profiles/dbg_filter.py 128 2>&1 | grep -v "warp  [0-9] sel\|warp  [5-7] epi\|warp  9 epi\|warp 1[01] epi" | tee gpurun_out/dbg_filter_d128.txt

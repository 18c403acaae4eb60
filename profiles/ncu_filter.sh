#!/bin/bash
# ncu full capture (with source-level stall sampling) of the filter kernel only
TAG=${1:-f}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter -s 2 -c 1 -f -o gpurun_out/prof_filter_$TAG \
    python profiles/run_filter.py 128 > /dev/null 2> gpurun_out/ncu_filter_$TAG.err
tail -3 gpurun_out/ncu_filter_$TAG.err; ls -la gpurun_out | tail -3

#!/bin/bash
# ncu full capture of the tensor-core filter kernel only
TAG=${1:-s}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter -s 2 -c 1 -f -o gpurun_out/prof_score_$TAG \
    python bench.py --steps 4 --warmup 3 --skip-cpu --skip-sweep > /dev/null 2> gpurun_out/ncu_score_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_|convert|merge" -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --skip-cpu --skip-sweep > /dev/null 2>&1
ls -la gpurun_out | tail -3

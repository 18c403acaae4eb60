#!/bin/bash
# One gpurun call: GPU tests, a clean bench line, the ncu launch list and full captures of the top kernels.
# usage (from the repo root, on the GPU box):  bash profiles/run_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --skip-cpu --skip-sweep > /dev/null 2> gpurun_out/ncu_launches_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_ -s 8 -c 4 -f -o gpurun_out/prof_bpr_$TAG \
    python bench.py --steps 4 --warmup 3 --skip-cpu --skip-score --skip-sweep > /dev/null 2> gpurun_out/ncu_bpr_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter -s 2 -c 1 -f -o gpurun_out/prof_score_$TAG \
    python bench.py --steps 4 --warmup 3 --skip-cpu --skip-sweep > /dev/null 2> gpurun_out/ncu_score_$TAG.err
ls -la gpurun_out | tail -12
SKIP=5 timeout 300 bash profiles/ncu_als.sh $TAG > /dev/null 2>&1
ls -la gpurun_out | tail -5

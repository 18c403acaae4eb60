#!/bin/bash
# round 2, final GPU call (1 GPU): build + smoke, the whole GPU suite, the bench line and the reference arm as the driver runs them,
# the launch list of the bench command, ncu --set full of the round's new kernels (blocked ALS solve, dataflow multi-step kernel)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_reference.json; cut -c1-300 gpurun_out/bench_final_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 3 --inner 32 --skip-cpu --skip-sweep > /dev/null 2> gpurun_out/ncu_launches_final.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_segment_kernel -s 5 -c 1 -f -o gpurun_out/prof_als_final \
    python profiles/als_probe.py 0.04 256 4096 144 0 > gpurun_out/ncu_als_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_flow_kernel -s 2 -c 1 -f -o gpurun_out/prof_flow_final \
    python profiles/run_flow.py > gpurun_out/ncu_flow_final.log 2>&1
tail -2 gpurun_out/ncu_flow_final.log

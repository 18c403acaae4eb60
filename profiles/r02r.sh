#!/bin/bash
# round 2, GPU call R (1 GPU): whole GPU suite + smoke, the bench line as the driver runs it, the reference arm, the launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_r02r.log; cat gpurun_out/pytest_r02r.log
timeout 900 python bench.py > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err; tail -c 1500 gpurun_out/bench_r02r.json; tail -5 gpurun_out/bench_r02r.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02r.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-sweep > /dev/null 2> gpurun_out/ncu_launches_r02r.err
tail -2 gpurun_out/ncu_launches_r02r.err

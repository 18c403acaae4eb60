#!/bin/bash
# round 2, GPU call A: the new BASELINE-config parity tests + the whole GPU suite, fold-0 BPR/VBPR acceptance,
# ncu capture of the HBM-streaming K1 configuration
mkdir -p gpurun_out/fold0
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_baseline.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r02a_baseline.log
cat gpurun_out/pytest_r02a_baseline.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline.py 2>&1 | tail -8 > gpurun_out/pytest_r02a.log
cat gpurun_out/pytest_r02a.log
timeout 900 python profiles/fold0_bpr.py data_fold0 gpurun_out/fold0 > gpurun_out/fold0_r02a.log 2>&1; tail -5 gpurun_out/fold0_r02a.log | cut -c1-600
timeout 300 python profiles/run_stream.py 10 > gpurun_out/stream_r02a.json 2>&1; cat gpurun_out/stream_r02a.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_ -s 9 -c 3 -f -o gpurun_out/prof_stream_r02a \
    python profiles/run_stream.py 3 > gpurun_out/ncu_stream_r02a.log 2>&1
ncu -i gpurun_out/prof_stream_r02a.ncu-rep --page raw --csv > gpurun_out/ncu_stream_r02a_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8

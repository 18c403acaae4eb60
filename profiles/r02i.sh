#!/bin/bash
# round 2, GPU call I (8 GPUs): fused DP + peer exchange correctness at 8 ranks, then bench.py at N=8 and N=4 as the driver runs it
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $RUN8 profiles/dp_ngpu.py > gpurun_out/dp_8gpu.json 2> gpurun_out/dp_8gpu.err; tail -c 1200 gpurun_out/dp_8gpu.json; tail -3 gpurun_out/dp_8gpu.err | cut -c1-200
timeout 300 $RUN8 profiles/score_ngpu.py > gpurun_out/score_8gpu.json 2> gpurun_out/score_8gpu.err; tail -c 900 gpurun_out/score_8gpu.json; tail -3 gpurun_out/score_8gpu.err | cut -c1-200
timeout 900 $RUN8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_r02i.json 2> gpurun_out/bench_8gpu_r02i.err; tail -c 2500 gpurun_out/bench_8gpu_r02i.json; tail -3 gpurun_out/bench_8gpu_r02i.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --skip-sweep > gpurun_out/bench_4gpu_r02i.json 2> gpurun_out/bench_4gpu_r02i.err; tail -c 1500 gpurun_out/bench_4gpu_r02i.json

#!/bin/bash
# round 2, GPU call T (8 GPUs): bench.py at N=8 exactly as the driver launches it (score path through the ring of sweep segments)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_r02t.json 2> gpurun_out/bench_8gpu_r02t.err
tail -c 2500 gpurun_out/bench_8gpu_r02t.json; tail -4 gpurun_out/bench_8gpu_r02t.err | cut -c1-300

#!/bin/bash
# 8 GPUs: ring of sweep segments after the whole-table seeding of first segments
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 profiles/ring_ngpu.py 18944 400 0 > gpurun_out/ring_8gpu_seeded.json 2> gpurun_out/ring_8gpu_seeded.err
cat gpurun_out/ring_8gpu_seeded.json | cut -c1-500; tail -2 gpurun_out/ring_8gpu_seeded.err | cut -c1-200

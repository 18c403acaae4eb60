#!/bin/bash
# 4 GPUs: ring of sweep segments with the whole-table seeding of first segments
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 profiles/ring_ngpu.py 18944 300 0 > gpurun_out/ring_4gpu_seeded.json 2> gpurun_out/ring_4gpu_seeded.err
cat gpurun_out/ring_4gpu_seeded.json | cut -c1-500

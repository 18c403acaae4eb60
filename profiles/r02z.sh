#!/bin/bash
# round 2, GPU call Z (8 GPUs): ring of sweep segments with skewed item shards (odd ranks, which end the sweeps and re-score, get
# shorter shards): tail_cost 0 .. 0.04 made it slower (first run), so -0.01 / -0.02 / -0.03 (even ranks, which start the sweeps, shorter), 400 batches each
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 profiles/ring_ngpu.py 18944 400 -0.01 -0.02 -0.03 > gpurun_out/ring_skew_8gpu.json 2> gpurun_out/ring_skew_8gpu.err
cat gpurun_out/ring_skew_8gpu.json | cut -c1-600; tail -3 gpurun_out/ring_skew_8gpu.err | cut -c1-300

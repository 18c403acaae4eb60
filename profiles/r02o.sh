#!/bin/bash
# round 2, GPU call O (2 GPUs): bench.py at N=2 as the driver runs it, score path through the ring of sweep segments (RingScorer)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r02o.json 2> gpurun_out/bench_2gpu_r02o.err
tail -c 3000 gpurun_out/bench_2gpu_r02o.json; tail -5 gpurun_out/bench_2gpu_r02o.err | cut -c1-300

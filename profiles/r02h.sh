#!/bin/bash
# round 2, GPU call H (2 GPUs): whole GPU suite on GPU 0, then bench.py under torchrun at N=2 (multi-GPU paths of the bench)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_r02h.log; cat gpurun_out/pytest_r02h.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --inner 128 > gpurun_out/bench_2gpu_r02h.json 2> gpurun_out/bench_2gpu_r02h.err
tail -c 5000 gpurun_out/bench_2gpu_r02h.json; tail -8 gpurun_out/bench_2gpu_r02h.err | cut -c1-300

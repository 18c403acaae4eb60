#!/bin/bash
# round 2, GPU call B (2 GPUs): fused data-parallel K1 exchange and the K3 peer exchange -- correctness + timing
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN profiles/dp_ngpu.py > gpurun_out/dp_${NG:-2}gpu.json 2> gpurun_out/dp_${NG:-2}gpu.err; tail -c 1500 gpurun_out/dp_${NG:-2}gpu.json; tail -5 gpurun_out/dp_${NG:-2}gpu.err
timeout 300 $RUN profiles/score_ngpu.py > gpurun_out/score_${NG:-2}gpu.json 2> gpurun_out/score_${NG:-2}gpu.err; tail -c 1500 gpurun_out/score_${NG:-2}gpu.json; tail -5 gpurun_out/score_${NG:-2}gpu.err

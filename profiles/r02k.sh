#!/bin/bash
# round 2, GPU call K (1 GPU): whole GPU suite after the VBPR pairwise mode and the Hogwild entry point
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_r02k.log; cat gpurun_out/pytest_r02k.log

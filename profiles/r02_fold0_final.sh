#!/bin/bash
# last check of the round (1 GPU): the BPR and VBPR blocks of train.py on the shipped fold 0 with the final code, GPU evaluator lists vs the oracle's
mkdir -p gpurun_out
timeout 420 python profiles/fold0_bpr.py data_fold0 gpurun_out/fold0 > gpurun_out/fold0_final.log 2>&1; tail -4 gpurun_out/fold0_final.log | cut -c1-700
cp gpurun_out/fold0/*.json gpurun_out/ 2>/dev/null; ls gpurun_out/fold0 | head

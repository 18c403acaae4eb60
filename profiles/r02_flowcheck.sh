#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bpr.py -q -x -k "dataflow or persistent or over_widths" 2>&1 | tail -3

#!/usr/bin/env python3
"""The K1 step on tables far beyond L2 (bench.py's HBM-streaming point: 6 M x 1 M rows, d=128, uniform triples, B=2^20),
alone, for ncu.  usage: python profiles/run_stream.py [seconds]   (a few milliseconds under ncu)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import bench  # noqa: E402

if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    print(json.dumps(bench.bpr_hbm_streaming(torch.device("cuda", 0), seconds=seconds)))

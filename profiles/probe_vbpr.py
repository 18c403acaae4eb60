#!/usr/bin/env python3
"""VBPR operating points at C3 (bench.vbpr_points) on their own.  usage: python profiles/probe_vbpr.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, bench, topkrec
dev = torch.device("cuda", 0)
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
for p in bench.vbpr_points(smp, dev, 64.0):
    print(p["batch_size"], "%.1f us/step" % p["us_per_step"], "%.3g triples/s" % p["triples_per_sec"], flush=True)

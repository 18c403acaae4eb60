#!/usr/bin/env python3
"""A few launches of the dataflow multi-step kernel on C2 (B = 1024, 256 steps per launch) for ncu.  usage: python profiles/run_flow.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import bench  # noqa: E402
import topkrec  # noqa: E402

dev = torch.device("cuda", 0)
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
st = {k: torch.from_numpy(v).to(dev) for k, v in bench.init_state_np(bench.N_USERS, bench.N_ITEMS, bench.D).items()}
cfg = topkrec.BprCfg(bench.N_USERS, bench.N_ITEMS, bench.D)
B, n_steps = 1024, 1024
ws = topkrec.bpr_workspace(cfg, B, dev)
loss = torch.zeros(n_steps, device=dev)
u, i, j = topkrec.bpr_sample(smp, 0, B * n_steps, dev)
for r in range(2):
    topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, n_steps, ws, loss)
torch.cuda.synchronize()
print("ok", float(loss.mean()))

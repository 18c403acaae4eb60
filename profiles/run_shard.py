#!/usr/bin/env python3
"""A few score + top-30 steps of 18 944 users against an item shard (what one rank of an N-GPU run computes), for ncu launch
lists.  usage: python profiles/run_shard.py [log2_items=17] [steps=3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, topkrec
ni = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
nb, D, k = 18944, 128, 30
g = torch.Generator(device=dev); g.manual_seed(4)
V = torch.randn(ni, D, device=dev, generator=g) * 0.1
U = [torch.randn(nb, D, device=dev, generator=g) * 0.1 for _ in range(2)]
ws = torch.empty(topkrec.lib().tkr_score_topk_tc_workspace_bytes(nb, ni, D, k, 0), dtype=torch.uint8, device=dev)
for t in range(steps + 1):
    topkrec.score_topk(U[t % 2], V, k, engine="tc", ws=ws, items_prepared=t > 0)
torch.cuda.synchronize()
print("ok")

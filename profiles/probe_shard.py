#!/usr/bin/env python3
"""K3 on one GPU as a function of the item-shard size (what a rank of an N-GPU item-sharded run sees) and of the seed
fraction of a sweep: ms per 18 944-user step, d=128, k=30.  usage: python profiles/probe_shard.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import topkrec  # noqa: E402

dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_seed_div.argtypes = [ctypes.c_int32]; L.tkr_debug_set_seed_div.restype = None
nb, D, k = 18944, 128, 30
g = torch.Generator(device=dev); g.manual_seed(4)
Vfull = torch.randn(1 << 20, D, device=dev, generator=g) * 0.1
U = [torch.randn(nb, D, device=dev, generator=g) * 0.1 for _ in range(4)]
out = []
for shift in (20, 19, 18, 17):
    ni = 1 << shift
    V = Vfull[:ni].contiguous()
    for div in (12, 8, 6, 4):
        L.tkr_debug_set_seed_div(div)
        ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nb, ni, D, k, 0), dtype=torch.uint8, device=dev)
        nfb = torch.zeros(1, dtype=torch.int32, device=dev)
        for t in range(3):
            topkrec.score_topk(U[t], V, k, engine="tc", ws=ws, items_prepared=t > 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(10):
            topkrec.score_topk(U[t % 4], V, k, engine="tc", ws=ws, items_prepared=True, n_fallback=nfb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append({"items": ni, "seed_div": div, "ms": ms, "tflops": 2.0 * nb * ni * D / (ms / 1e3) / 1e12, "fallback_rows": int(nfb.item())})
        print(out[-1], flush=True)
L.tkr_debug_set_seed_div(12)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_shard.json"), "w"), indent=1)

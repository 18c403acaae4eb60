#!/usr/bin/env python3
"""K3 on item shards: step time and rows sent to the exact fallback for the seed rank / compaction trigger of the filter
(tkr_debug_set_filter_tuning).  18 944 users, d=128, k=30.  usage: python profiles/probe_filter_tuning.py"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, topkrec, bench
dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_filter_tuning.argtypes = [ctypes.c_int32, ctypes.c_int32]; L.tkr_debug_set_filter_tuning.restype = None
nb, D, k = 18944, 128, 30
g = torch.Generator(device=dev); g.manual_seed(4)
Vfull = torch.randn(1 << 20, D, device=dev, generator=g) * 0.1
U = [torch.randn(nb, D, device=dev, generator=g) * 0.1 for _ in range(4)]
out = []
for shift in (17, 18, 19, 20):
    ni = 1 << shift
    V = Vfull[:ni].contiguous()
    ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nb, ni, D, k, 0), dtype=torch.uint8, device=dev)
    for rank, cap in ((4, 128), (3, 128), (4, 96), (3, 96), (3, 80)):
        L.tkr_debug_set_filter_tuning(rank, cap)
        nfb = torch.zeros(1, dtype=torch.int32, device=dev)
        fb = 0
        topkrec.score_topk(U[0], V, k, engine="tc", ws=ws)
        st = {"t": 0}

        def run():
            st["t"] += 1
            topkrec.score_topk(U[st["t"] % 4], V, k, engine="tc", ws=ws, items_prepared=True, n_fallback=nfb)
        ms = bench.device_time_ms(run, 12, 4)
        for t in range(4):
            run(); fb += int(nfb.item())
        out.append({"items": ni, "seed_rank": rank, "cap_trigger": cap, "ms": ms, "fallback_rows_4_batches": fb})
        print(out[-1], flush=True)
L.tkr_debug_set_filter_tuning(0, 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_filter_tuning.json"), "w"), indent=1)

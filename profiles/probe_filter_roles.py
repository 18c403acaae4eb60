#!/usr/bin/env python3
"""Cycle accounting of the K3 filter kernel per warp role (tkr_debug_set_filter_counters) for a long sweep (2^20 items) and the
short sweep one rank of an 8-GPU item-sharded run sees (2^17 items).  usage: python profiles/probe_filter_roles.py"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, topkrec
dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_filter_counters.argtypes = [ctypes.c_void_p]; L.tkr_debug_set_filter_counters.restype = None
L.tkr_debug_set_filter_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_filter_mode.restype = None
nb, D, k, FW = 18944, 128, 30, 22
g = torch.Generator(device=dev); g.manual_seed(4)
Vfull = torch.randn(1 << 20, D, device=dev, generator=g) * 0.1
U = torch.randn(nb, D, device=dev, generator=g) * 0.1
for shift in (20, 17):
    ni = 1 << shift
    V = Vfull[:ni].contiguous()
    ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nb, ni, D, k, 0), dtype=torch.uint8, device=dev)
    topkrec.score_topk(U, V, k, engine="tc", ws=ws)
    n_ctas = 148
    dbg = torch.zeros(n_ctas * FW * 4, dtype=torch.int64, device=dev)
    L.tkr_debug_set_filter_counters(dbg.data_ptr()); L.tkr_debug_set_filter_mode(1)
    topkrec.score_topk(U, V, k, engine="tc", ws=ws, items_prepared=True)
    torch.cuda.synchronize()
    L.tkr_debug_set_filter_counters(None)
    c = dbg.cpu().numpy().reshape(n_ctas, FW, 4).astype(np.float64)
    sel, epi, tma, mma = c[:, 0:4], c[:, 4:20], c[:, 20], c[0::2, 21]
    tiles = ni // 256 + (ni // 256) // (8 if ni // 256 < 2048 else 12)
    print("items 2^%d: tiles/sweep %d" % (shift, tiles))
    print("  MMA thread   total %.0f kcyc  wait-for-stage %.0f kcyc  wall %.0f us  -> %.0f cyc/tile" % (mma[:, 0].mean() / 1e3, mma[:, 2].mean() / 1e3, mma[:, 3].mean() / 1e3, mma[:, 0].mean() / tiles))
    print("  TMA thread   total %.0f kcyc  wait-for-empty %.0f kcyc" % (tma[:, 0].mean() / 1e3, tma[:, 1].mean() / 1e3))
    print("  epilogue     scan %.0f kcyc  wait-for-tile %.0f kcyc  drain+handback %.0f kcyc (tmem load %.0f)" % (epi[:, :, 0].mean() / 1e3, epi[:, :, 1].mean() / 1e3, epi[:, :, 2].mean() / 1e3, epi[:, :, 3].mean() / 1e3))
    print("  selection    busy %.0f kcyc  idle %.0f kcyc  blocks %.0f  compaction %.0f kcyc" % (sel[:, :, 0].mean() / 1e3, sel[:, :, 1].mean() / 1e3, sel[:, :, 2].mean(), sel[:, :, 3].mean() / 1e3))

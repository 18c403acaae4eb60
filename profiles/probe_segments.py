#!/usr/bin/env python3
"""Where a ring slot's time goes, on ONE GPU: one sweep of 18 944 users x 1 M items (d=128, k=30) cut into G segments with
tkr_score_topk_tc_segment (one workspace per shard so that every shard's BF16 table stays prepared, as on its own rank), each
segment timed with CUDA events; against the unsegmented sweep.  usage: python profiles/probe_segments.py [G=8] [reps=20] [seed_div]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import topkrec  # noqa: E402
from topkrec import dist as tdist  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
nb, NI, D, k = 18944, 1 << 20, 128, 30
L = topkrec.lib()
if len(sys.argv) > 3:
    L.tkr_debug_set_seed_div(int(sys.argv[3]))
g = torch.Generator(device=dev); g.manual_seed(4)
Vfull = torch.randn(NI, D, device=dev, generator=g) * 0.1
U = [torch.randn(nb, D, device=dev, generator=g) * 0.1 for _ in range(4)]
bounds = tdist.shard_bounds(NI, G)
Vs = [Vfull[b:e].contiguous() for b, e in bounds]
wss = [torch.empty(L.tkr_score_topk_tc_segment_workspace_bytes(nb, e - b, NI, D, k, 0), dtype=torch.uint8, device=dev) for b, e in bounds]
state = torch.empty(L.tkr_score_topk_tc_state_bytes(nb), dtype=torch.uint8, device=dev)
out = (torch.empty((nb, k), dtype=torch.int32, device=dev), torch.empty((nb, k), dtype=torch.float32, device=dev))
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(G + 1)] for _ in range(reps)]


def sweep(r, prepared, record):
    for s in range(G):
        if record: ev[r][s].record()
        topkrec.score_topk_segment(U[r % 4], Vs[s], k, bounds[s][0], state, s == 0, s == G - 1, V_full=Vfull, out=out if s == G - 1 else None,
                                   ws=wss[s], items_prepared=prepared)
    if record: ev[r][G].record()


sweep(0, False, False); sweep(1, True, False)
torch.cuda.synchronize()
for r in range(reps):
    sweep(r, True, True)
torch.cuda.synchronize()
seg_ms = [sum(ev[r][s].elapsed_time(ev[r][s + 1]) for r in range(reps)) / reps for s in range(G)]
ws1 = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nb, NI, D, k, 0), dtype=torch.uint8, device=dev)
topkrec.score_topk(U[0], Vfull, k, engine="tc", ws=ws1, items_prepared=False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(reps):
    wi, wsc = topkrec.score_topk(U[r % 4], Vfull, k, engine="tc", ws=ws1, items_prepared=True)
e1.record(); torch.cuda.synchronize()
whole = e0.elapsed_time(e1) / reps
same = bool(torch.equal(out[0], wi) and torch.equal(out[1].view(torch.int32), wsc.view(torch.int32)))
print(json.dumps({"seed_div": int(sys.argv[3]) if len(sys.argv) > 3 else 12, "segments": G, "segment_ms": [round(x, 4) for x in seg_ms], "sum_ms": round(sum(seg_ms), 4), "whole_sweep_ms": round(whole, 4),
                  "whole_over_G_ms": round(whole / G, 4), "last_batch_equals_whole_sweep_bitwise": same}))

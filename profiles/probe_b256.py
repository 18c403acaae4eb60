#!/usr/bin/env python3
"""K1 at the reference's batch size (256) on C2 and on the fold-0 shape: us per step of the persistent cluster kernel vs
the two-launch route (tkr_debug_set_persist_mode), explicit triples and fused sampler.  usage: python profiles/probe_b256.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import bench  # noqa: E402
import topkrec  # noqa: E402

dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_persist_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_persist_mode.restype = None
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
out = []
for (nu, ni, d) in ((70000, 10000, 128), (70000, 10000, 50)):
    st = {k: torch.from_numpy(v).to(dev) for k, v in bench.init_state_np(nu, ni, d).items()}
    cfg = topkrec.BprCfg(nu, ni, d)
    for B in (64, 256, 512, 1024):
        n_steps = 2048
        ws = topkrec.bpr_workspace(cfg, B, dev)
        loss = torch.zeros(n_steps, device=dev)
        trip = topkrec.bpr_sample(smp, 0, B * n_steps, dev)
        for mode, name in ((2, "dataflow"), (1, "persistent"), (0, "two_kernel")):
            L.tkr_debug_set_persist_mode(mode)
            for fused in (False, True):
                t3 = (None, None, None) if fused else trip

                def run(r):
                    topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], *t3, B, n_steps, ws, loss,
                                     sampler=smp if fused else None, first_draw=r * B * n_steps)
                run(0); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for r in range(3):
                    run(1 + r)
                e1.record(); torch.cuda.synchronize()
                us = 1e3 * e0.elapsed_time(e1) / (3 * n_steps)
                out.append({"d": d, "batch": B, "route": name, "fused_sampler": fused, "us_per_step": us, "triples_per_s": B / (us / 1e6)})
                print(out[-1], flush=True)
L.tkr_debug_set_persist_mode(-1)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_b256.json"), "w"), indent=1)

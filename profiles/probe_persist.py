#!/usr/bin/env python3
"""Where a step of the persistent B=256 kernel spends its cycles (warp 0's clock64 deltas per phase)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, bench, topkrec
dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_persist_counters.argtypes = [ctypes.c_void_p]; L.tkr_debug_set_persist_counters.restype = None
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
for (nu, ni, d, B) in ((70000, 10000, 128, 256), (70000, 10000, 128, 64), (2000, 1000, 128, 256), (70000, 10000, 50, 256)):
    st = {k: torch.from_numpy(v).to(dev) for k, v in bench.init_state_np(nu, ni, d).items()}
    cfg = topkrec.BprCfg(nu, ni, d)
    n_steps = 2048
    ws = topkrec.bpr_workspace(cfg, B, dev)
    u = torch.randint(0, nu, (B * n_steps,), device=dev, dtype=torch.int32); i = torch.randint(0, ni, (B * n_steps,), device=dev, dtype=torch.int32)
    j = torch.randint(0, ni, (B * n_steps,), device=dev, dtype=torch.int32)
    for with_loss in (True, False):
        loss = torch.zeros(n_steps, device=dev) if with_loss else None
        dbg = torch.zeros(8, dtype=torch.int64, device=dev)
        topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, n_steps, ws, loss)
        torch.cuda.synchronize()
        L.tkr_debug_set_persist_counters(dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, n_steps, ws, loss)
        e1.record(); torch.cuda.synchronize()
        L.tkr_debug_set_persist_counters(None)
        c = dbg.cpu().numpy()
        n = max(1, int(c[5]))
        print(dict(nu=nu, ni=ni, d=d, B=B, loss=with_loss, us_per_step=1e3 * e0.elapsed_time(e1) / n_steps,
                   cycles=dict(gather_grad=int(c[0] // n), slot_prefetch=int(c[1] // n), barrier1=int(c[2] // n), update=int(c[3] // n), barrier2=int(c[4] // n))), flush=True)

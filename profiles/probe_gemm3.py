#!/usr/bin/env python3
"""3xTF32 content GEMMs of VBPR at C3: accuracy of the projection vs fp64 and step time for the GEMM variants
(tkr_debug_set_gemm3_flags: 0 = fused N=2NP product, 2 = three products, 1 = raw A tile kept as the hi part)."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, bench, topkrec
dev = torch.device("cuda", 0)
L = topkrec.lib()
L.tkr_debug_set_gemm3_flags.argtypes = [ctypes.c_int32]; L.tkr_debug_set_gemm3_flags.restype = None
L.tkr_debug_set_vbpr_tc_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_vbpr_tc_mode.restype = None
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
rng = np.random.default_rng(7)
ni, dF, k = bench.N_ITEMS, 4096, 128
h = k // 2
Fh = np.abs(rng.standard_normal((ni, dF), dtype=np.float32)); Fh /= np.linalg.norm(Fh, axis=1, keepdims=True)
E = (0.05 * rng.standard_normal((dF, h))).astype(np.float32); c = (0.05 * rng.standard_normal(dF)).astype(np.float32)
P = Fh.astype(np.float64) @ E.astype(np.float64)
F = torch.from_numpy(Fh).to(dev)
cfg = topkrec.VbprCfg(bench.N_USERS, ni, k, dF)
out = []
for flags, tc in ((0, -1), (1, -1)):
    L.tkr_debug_set_gemm3_flags(flags); L.tkr_debug_set_vbpr_tc_mode(tc)
    g = torch.Generator(device=dev); g.manual_seed(2)
    st = {"U": torch.randn(bench.N_USERS, k, device=dev, generator=g) * 0.01, "V": torch.zeros(ni, k, device=dev),
          "rb": torch.zeros(ni, device=dev), "bsum": torch.zeros(ni, device=dev), "E": torch.from_numpy(E).to(dev), "c": torch.from_numpy(c).to(dev)}
    st["V"][:, :h] = torch.randn(ni, h, device=dev, generator=g) * 0.01
    for n, m in (("U", "msU"), ("V", "msV"), ("rb", "msrb"), ("E", "msE"), ("c", "msc")):
        st[m] = torch.ones_like(st[n])
    rec = {"gemm3_flags": flags, "route": "tcgen05 3xTF32" if tc else "fp32 CUDA cores"}
    for B, n_steps in ((1 << 16, 16), (1 << 20, 4)):
        ws = topkrec.vbpr_workspace(cfg, B, dev)
        topkrec.vbpr_set_hot_items(cfg, B, ws, topkrec.popular_items(smp.pos_idx, ni))
        z = torch.zeros(B, dtype=torch.int32, device=dev)
        if B == 1 << 16:
            topkrec.vbpr_step(cfg, st, F, z, z, z, B, 0, ws, None)
            rec["projection_rel_err_vs_fp64"] = float(np.abs(st["V"].cpu().numpy()[:, h:] - P).max() / np.abs(P).max())
        loss = torch.zeros(n_steps, device=dev)
        s = {"r": 0}

        def run():
            s["r"] += 1
            topkrec.vbpr_step(cfg, st, F, None, None, None, B, n_steps, ws, loss, sampler=smp, first_draw=s["r"] * B * n_steps)
        one = bench.device_time_ms(run, 2, 2)
        ms = bench.device_time_ms(run, int(max(3, min(100, 500 / one))), 0) / n_steps
        rec["us_per_step_B%d" % B] = 1e3 * ms
        del ws
    print(rec, flush=True)
    out.append(rec)
L.tkr_debug_set_gemm3_flags(0); L.tkr_debug_set_vbpr_tc_mode(-1)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_gemm3.json"), "w"), indent=1)

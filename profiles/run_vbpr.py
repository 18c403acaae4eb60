#!/usr/bin/env python3
"""A few VBPR steps at C3 (70k x 10k, 4096-d features, k=128) for ncu launch lists.  usage: python profiles/run_vbpr.py [log2_batch] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch, bench, topkrec
B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
tr_users, indptr, pos_idx = bench.synth_interactions()
smp = topkrec.Sampler(tr_users, indptr, pos_idx, bench.N_ITEMS, seed=123, device=dev)
g = torch.Generator(device=dev); g.manual_seed(2)
d_feat, k = 4096, 128
h = k // 2
F = torch.randn(bench.N_ITEMS, d_feat, device=dev, generator=g).abs_(); F /= F.norm(dim=1, keepdim=True)
cfg = topkrec.VbprCfg(bench.N_USERS, bench.N_ITEMS, k, d_feat)
st = {"U": torch.randn(bench.N_USERS, k, device=dev, generator=g) * 0.01, "V": torch.zeros(bench.N_ITEMS, k, device=dev),
      "rb": torch.zeros(bench.N_ITEMS, device=dev), "bsum": torch.zeros(bench.N_ITEMS, device=dev),
      "E": torch.full((d_feat, h), 2.0 / (d_feat * k), device=dev), "c": torch.zeros(d_feat, device=dev)}
st["V"][:, :h] = torch.randn(bench.N_ITEMS, h, device=dev, generator=g) * 0.01
for n, m in (("U", "msU"), ("V", "msV"), ("rb", "msrb"), ("E", "msE"), ("c", "msc")):
    st[m] = torch.ones_like(st[n])
ws = topkrec.vbpr_workspace(cfg, B, dev)
topkrec.vbpr_set_hot_items(cfg, B, ws, topkrec.popular_items(smp.pos_idx, bench.N_ITEMS))
loss = torch.zeros(steps, device=dev)
for r in range(2):
    topkrec.vbpr_step(cfg, st, F, None, None, None, B, steps, ws, loss, sampler=smp, first_draw=r * B * steps)
torch.cuda.synchronize()
print("ok", loss.tolist())

import sys
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import numpy as np, torch, topkrec
def rel(a, b): return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
nu, ni, d, B = 70000, 10000, 128, 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = {"U": torch.randn(nu, d, device="cuda", generator=g) * 0.01, "V": torch.randn(ni, d, device="cuda", generator=g) * 0.01, "b": torch.zeros(ni, device="cuda")}
a.update({"ms" + k: torch.ones_like(v) for k, v in list(a.items())})
rng = np.random.default_rng(0)
u = torch.from_numpy(rng.integers(0, nu // 2, B).astype(np.int32)).cuda()
p = 1.0 / np.arange(1, ni + 1); p /= p.sum()
i = torch.from_numpy(rng.choice(ni, B, p=p).astype(np.int32)).cuda()
j = torch.from_numpy(rng.integers(0, ni, B).astype(np.int32)).cuda()
cfg = topkrec.BprCfg(nu, ni, d)
ws = topkrec.bpr_workspace(cfg, B)
outs = []
for hot in (False, False, True, True):
    topkrec.bpr_set_hot_items(cfg, B, ws, topkrec.popular_items(i, ni) if hot else [])
    b = {k: v.clone() for k, v in a.items()}
    topkrec.bpr_step(cfg, b["U"], b["V"], b["b"], b["msU"], b["msV"], b["msb"], u, i, j, B, 1, ws, None)
    outs.append({k: v.cpu().numpy() for k, v in b.items()})
for x in range(1, 4):
    print('run', x, 'vs 0:', {n: float(rel(outs[x][n], outs[0][n])) for n in outs[0]})
d_ = np.abs(outs[1]["msV"] - outs[0]["msV"]); r = np.unravel_index(d_.argmax(), d_.shape); print('worst msV entry', r, outs[0]["msV"][r], outs[1]["msV"][r], 'V', outs[0]["V"][r], outs[1]["V"][r])

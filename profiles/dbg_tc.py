import sys, os, time
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import numpy as np, torch, topkrec
from oracle import topk_ref
torch.cuda.init()
def run(nu, ni, d, k, bias=False, rated=0, seed=0, ties=False):
    rng = np.random.default_rng(seed)
    U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
    V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
    if ties:
        V[rng.integers(0, ni, ni // 10)] = V[0]; V[rng.integers(0, ni, ni // 20)] = 0
    b = (0.05 * rng.standard_normal(ni)).astype(np.float32) if bias else None
    indptr = idx = None
    if rated:
        cnt = rng.integers(0, rated + 1, nu); indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum(cnt)
        idx = np.concatenate([np.sort(rng.choice(ni, c, replace=False)) for c in cnt] + [np.zeros(0, np.int64)]).astype(np.int32)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()
    nf = torch.zeros(1, dtype=torch.int32, device='cuda')
    t0 = time.time()
    gi, gs = topkrec.score_topk(t(U), t(V), k, t(b), t(indptr), t(idx), engine='tc', n_fallback=nf)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ri, rs = topk_ref.score_topk(U, V, k, b, indptr, idx)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    bad = np.nonzero((gi != ri).any(1))[0]
    print('nu=%d ni=%d d=%d k=%d bias=%s rated=%d ties=%s: idx_equal=%s score_equal=%s bad_rows=%d fallback_rows=%d  %.1f ms' % (
        nu, ni, d, k, bias, rated, ties, np.array_equal(gi, ri), np.array_equal(gs.view(np.uint32), rs.view(np.uint32)), len(bad), nf.item(), dt * 1e3), flush=True)
    if len(bad):
        r = bad[0]; print(' row', r, '\n got', gi[r][:12], gs[r][:6], '\n ref', ri[r][:12], rs[r][:6])
run(128, 256, 64, 10)
run(128, 1024, 128, 30)
run(300, 5000, 128, 30)
run(300, 5000, 50, 30, bias=True, rated=40)
run(1000, 20000, 128, 30, rated=64, ties=True)
run(200, 3000, 250, 30, bias=True)
run(19000, 100000, 128, 30)

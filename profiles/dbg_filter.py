"""cycle counters of the filter kernel roles at the bench shape (profiling aid)"""
import sys, ctypes
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import numpy as np, torch, topkrec
nu, ni, d, k = 18944, 1 << 20, 128, 30
g = torch.Generator(device='cuda'); g.manual_seed(4)
V = torch.randn(ni, d, device='cuda', generator=g) * 0.1
U = torch.randn(nu, d, device='cuda', generator=g) * 0.1
L = topkrec.lib()
L.tkr_debug_set_filter_counters.argtypes = [ctypes.c_void_p]; L.tkr_debug_set_filter_counters.restype = None
dbg = torch.zeros(148 * 10 * 4, dtype=torch.int64, device='cuda')
ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 0), dtype=torch.uint8, device='cuda')
for it in range(3):
    topkrec.score_topk(U, V, k, engine='tc', ws=ws)
L.tkr_debug_set_filter_counters(dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws); e1.record(); torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1))
x = dbg.cpu().numpy().reshape(148, 10, 4).astype(np.float64)
tiles = 4096
print('per-tile cycles (mean over CTAs):')
for w in range(10):
    role = 'epi' if w < 8 else ('tma' if w == 8 else 'mma')
    print(' warp %d %s c0 %.0f c1 %.0f c2 %.0f c3 %.0f   (epi: scan, wait_tfull, tmem_ld, compact)' % (w, role, x[:, w, 0].mean() / tiles, x[:, w, 1].mean() / tiles, x[:, w, 2].mean() / tiles, x[:, w, 3].mean() / tiles))
print('cta spread of total cycles: min %.3g max %.3g' % (x[:, 9, 0].min(), x[:, 9, 0].max()))

# experiment: no row ever collects (tau = +inf): pure tcgen05.ld + max-tree epilogue
dbg.zero_(); dbg[0] = -12345
e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws); e1.record(); torch.cuda.synchronize()
print('NO-HIT step ms', e0.elapsed_time(e1))
x = dbg.cpu().numpy().reshape(148, 10, 4).astype(np.float64)
for w in (0, 4, 8, 9):
    print(' warp %d c0 %.0f c1 %.0f c2 %.0f c3 %.0f' % (w, x[1:, w, 0].mean() / tiles, x[1:, w, 1].mean() / tiles, x[1:, w, 2].mean() / tiles, x[1:, w, 3].mean() / tiles))

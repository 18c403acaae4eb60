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
dbg = torch.zeros(148 * 14 * 4, dtype=torch.int64, device='cuda')
ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 0), dtype=torch.uint8, device='cuda')
for it in range(3):
    topkrec.score_topk(U, V, k, engine='tc', ws=ws)
L.tkr_debug_set_filter_counters(dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws); e1.record(); torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1))
x = dbg.cpu().numpy().reshape(148, 14, 4).astype(np.float64)
tiles = 4096
print('per-tile cycles (mean over CTAs):')
for w in range(14):
    role = 'epi(scan,wait_tfull,tmem_ld)' if w < 8 else ('sel(busy,idle,events,compact)' if w < 12 else ('tma(total,wait_empty)' if w == 12 else 'mma(total,wait_tempty,wait_full)'))
    print(' warp %d %s c0 %.0f c1 %.0f c2 %.0f c3 %.0f' % (w, role, x[:, w, 0].mean() / tiles, x[:, w, 1].mean() / tiles, x[:, w, 2].mean() / tiles, x[:, w, 3].mean() / tiles))
print('cta spread of total cycles: min %.3g max %.3g' % (x[:, 13, 0].min(), x[:, 13, 0].max()))


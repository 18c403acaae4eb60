"""cycle counters of the filter kernel roles at the bench shape (profiling aid); modes 2/3 are ceiling probes"""
import sys, ctypes
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import numpy as np, torch, topkrec
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nu, ni, k = 18944, 1 << 20, 30
g = torch.Generator(device='cuda'); g.manual_seed(4)
V = torch.randn(ni, d, device='cuda', generator=g) * 0.1
U = torch.randn(nu, d, device='cuda', generator=g) * 0.1
L = topkrec.lib()
L.tkr_debug_set_filter_counters.argtypes = [ctypes.c_void_p]; L.tkr_debug_set_filter_counters.restype = None
L.tkr_debug_set_filter_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_filter_mode.restype = None
L.tkr_debug_filter_max_pairs.argtypes = [ctypes.c_int32]; L.tkr_debug_filter_max_pairs.restype = ctypes.c_int32
print('resident CTA pairs:', L.tkr_debug_filter_max_pairs(d))
dbg = torch.zeros(148 * 22 * 4, dtype=torch.int64, device='cuda')
ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 0), dtype=torch.uint8, device='cuda')
for it in range(3):
    topkrec.score_topk(U, V, k, engine='tc', ws=ws)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws); e1.record(); torch.cuda.synchronize()
print('d=%d product step ms %.3f' % (d, e0.elapsed_time(e1)))
tiles = 4096 + 64
for mode in (1, 2, 3, 4, 5):
    dbg.zero_()
    L.tkr_debug_set_filter_counters(dbg.data_ptr()); L.tkr_debug_set_filter_mode(mode)
    e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws); e1.record(); torch.cuda.synchronize()
    L.tkr_debug_set_filter_counters(None)
    x = dbg.cpu().numpy().reshape(148, 22, 4).astype(np.float64)
    print('mode %d (%s): call ms %.3f; MMA-warp cycles per tile %.0f (ideal %d)' % (
        mode, {1: 'normal+counters', 2: 'no TMEM reads', 3: 'TMEM drain, no scan', 4: 'no TMEM reads, no V loads', 5: 'full scan, thresholds at +inf (no hand-offs)'}[mode], e0.elapsed_time(e1), x[0::2, 21, 0].mean() / tiles, 128 * ((d + 63) // 64) * 4))
    for w in (0, 4, 8, 12, 16, 20, 21):
        role = 'sel(busy,idle,events,compact)' if w < 4 else ('epi(scan,wait_tfull,drain+handback,tmem_ld)' if w < 20 else ('tma(total,wait_empty)' if w == 20 else 'mma(total,-,wait_full,ns)'))
        print('   warp %2d %-36s %7.0f %7.0f %7.0f %7.0f' % (w, role, x[:, w, 0].mean() / tiles, x[:, w, 1].mean() / tiles, x[:, w, 2].mean() / tiles, x[:, w, 3].mean() / tiles))
    lead = x[0::2, 21]
    print('   MMA loop: %.3f ms, SM clock %.0f MHz, %.0f TFLOP/s over the loop' % (lead[:, 3].mean() * 1e-6, lead[:, 0].mean() / lead[:, 3].mean() * 1e3, 2.0 * nu * 256 * tiles * d / (lead[:, 3].mean() * 1e-9) / 1e12))
    print('   cta spread of MMA total cycles: min %.4g max %.4g' % (x[0::2, 21, 0].min(), x[0::2, 21, 0].max()))

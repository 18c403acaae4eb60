"""step time of the tensor-core score+top-k at the bench shape vs the seed fraction (tuning aid)"""
import sys, ctypes
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import torch, topkrec
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nu, ni, k = 18944, 1 << 20, 30
g = torch.Generator(device='cuda'); g.manual_seed(4)
V = torch.randn(ni, d, device='cuda', generator=g) * 0.1
U = torch.randn(nu, d, device='cuda', generator=g) * 0.1
L = topkrec.lib()
L.tkr_debug_set_seed_div.argtypes = [ctypes.c_int32]; L.tkr_debug_set_seed_div.restype = None
ws = torch.empty(L.tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 0), dtype=torch.uint8, device='cuda')
nf = torch.zeros(1, dtype=torch.int32, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for div in [int(a) for a in sys.argv[2:]] or [64, 32, 16, 8, 6, 4]:
    L.tkr_debug_set_seed_div(div)
    topkrec.score_topk(U, V, k, engine='tc', ws=ws)
    ts = []
    for it in range(5):
        e0.record(); topkrec.score_topk(U, V, k, engine='tc', ws=ws, items_prepared=True, n_fallback=nf); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print('d=%d seed 1/%d: %.3f ms/step (items prepared) = %.0f TFLOP/s, %.2f M users/s, fallback rows %d' % (d, div, ms, 2.0 * nu * ni * d / ms / 1e9, nu / ms / 1e3, nf.item()), flush=True)

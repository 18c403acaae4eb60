"""three calls of the tensor-core score+top-k at the bench shape (target of the ncu captures)"""
import sys
sys.path.insert(0, 'top-k-rec_b200'); sys.path.insert(0, '.')
import torch, topkrec
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nu, ni, k = 18944, 1 << 20, 30
g = torch.Generator(device='cuda'); g.manual_seed(4)
V = torch.randn(ni, d, device='cuda', generator=g) * 0.1
U = torch.randn(nu, d, device='cuda', generator=g) * 0.1
ws = torch.empty(topkrec.lib().tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, 0), dtype=torch.uint8, device='cuda')
for it in range(3):
    topkrec.score_topk(U, V, k, engine='tc', ws=ws, items_prepared=it > 0)
torch.cuda.synchronize()

#!/usr/bin/env python3
"""Item-sharded score + top-30 on N GPUs (torchrun): the peer-memory exchange (topkrec.dist.ShardedScorer) against the
unsharded single-GPU result (bit-identical lists and scores required, with and without a rated mask) and against round 1's
all-gather route, then device-timed pipelined steps of both on the fixed 18 944-user x 1 M-item batch (d=128, k=30).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/score_ngpu.py [users] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import topkrec  # noqa: E402
from topkrec import dist as tdist  # noqa: E402


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    NI, D, k = 1 << 20, 128, 30
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev); g.manual_seed(4)
    Vfull = torch.randn(NI, D, device=dev, generator=g) * 0.1          # the same table on every rank; each keeps its shard
    beg, end = tdist.shard_bounds(NI, world)[rank]
    V = Vfull[beg:end].contiguous()
    gu = torch.Generator(device=dev); gu.manual_seed(3)
    Ub = [torch.randn(nb, D, device=dev, generator=gu) * 0.1 for _ in range(4)]
    ri = torch.sort(torch.randint(0, NI, (nb, 64), device=dev, generator=gu, dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
    rp = torch.arange(0, (nb + 1) * 64, 64, device=dev, dtype=torch.int64)
    out = {"world": world, "users_per_step": nb, "items": NI, "d": D, "k": k}
    sc = tdist.ShardedScorer(end - beg, D, k, nb, beg, device=dev)
    # ---- correctness: own slice == the unsharded engine on the same rows (3 batches: exercises both parities + reuse)
    ok = True
    for t, masked in ((0, False), (1, True), (2, False), (3, True)):
        oi, osc = sc.submit(Ub[t], V, None, rp if masked else None, ri if masked else None)
        sc.wait()
        b0, b1 = sc.rows_of(nb)
        wi, wsc = topkrec.score_topk(Ub[t][b0:b1].contiguous(), Vfull, k, None, rp[b0:b1 + 1].contiguous() - rp[b0] if masked else None,
                                     ri[rp[b0]:rp[b1]].contiguous() if masked else None, engine="tc")
        ok = ok and bool(torch.equal(oi, wi)) and bool(torch.equal(osc.view(torch.int32), wsc.view(torch.int32)))
    flag = torch.tensor([int(ok)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["sharded_equals_unsharded_bitwise_all_ranks"] = bool(flag.item())
    del Vfull
    torch.cuda.empty_cache()

    def timed(fn, sync):
        for t in range(4):
            fn(t)
        sync(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(K):
            fn(t)
        sync()
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms = timed(lambda t: sc.submit(Ub[t % 4], V), sc.wait)
    out["peer_exchange"] = {"ms_per_step": ms, "users_per_s": nb / (ms / 1e3)}
    wsb = torch.empty(topkrec.lib().tkr_score_topk_tc_workspace_bytes(nb, end - beg, D, k, 0), dtype=torch.uint8, device=dev)
    st = {"p": False}

    def old(t):
        tdist.sharded_score_topk(Ub[t % 4], V, k, beg, engine="tc", ws=wsb, items_prepared=st["p"])
        st["p"] = True
    ms = timed(old, lambda: None)
    out["nccl_all_gather_round1"] = {"ms_per_step": ms, "users_per_s": nb / (ms / 1e3)}
    # the per-rank kernels alone (no exchange): what the shard costs
    ms = timed(lambda t: topkrec.score_topk(Ub[t % 4], V, k, col_offset=beg, engine="tc", ws=wsb, items_prepared=True), lambda: None)
    out["local_scoring_only"] = {"ms_per_step": ms}
    sc.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

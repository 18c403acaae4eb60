#!/usr/bin/env python3
"""Item-sharded score + top-30 as a ring of sweep segments on N GPUs (torchrun): topkrec.dist.RingScorer against the unsharded
engine (bit-identical lists and score bits required on the last rank, with and without a rated mask), then device-timed
pipelined steps on the fixed 18 944-user x 1 M-item batch (d=128, k=30).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/ring_ngpu.py [users] [steps] [tail_cost ...]
(tail_cost: dist.ring_shard_bounds -- 0 = equal shards; several values = one ring each, timed one after the other)"""
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
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    tails = [float(a) for a in sys.argv[3:]] or [0.0]
    NI, D, k = 1 << 20, 128, 30
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev); g.manual_seed(4)
    Vfull = torch.randn(NI, D, device=dev, generator=g) * 0.1
    gu = torch.Generator(device=dev); gu.manual_seed(3)
    Ub = [torch.randn(nb, D, device=dev, generator=gu) * 0.1 for _ in range(4)]
    ri = torch.sort(torch.randint(0, NI, (nb, 64), device=dev, generator=gu, dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
    rp = torch.arange(0, (nb + 1) * 64, 64, device=dev, dtype=torch.int64)
    for tail in tails:
        bounds = tdist.ring_shard_bounds(NI, world, tail)
        beg, end = bounds[rank]
        V = Vfull[beg:end].contiguous()
        out = {"world": world, "users_per_step": nb, "items": NI, "d": D, "k": k, "tail_cost": tail, "shard_items": [e - b for b, e in bounds]}
        ring = tdist.RingScorer(V, D, k, nb, beg, Vfull, device=dev)
        # ---- correctness: 2 G + 3 batches (every owner gets some), odd ones masked; the owner compares with the unsharded engine
        T = 2 * world + 3
        batches = [Ub[t % 4] for t in range(T)]
        rated = [(rp, ri) if t % 2 else (None, None) for t in range(T)]
        bad = []

        def check(t, idx, score):
            wi, wsc = topkrec.score_topk(batches[t], Vfull, k, None, rated[t][0], rated[t][1], engine="tc")
            if not (torch.equal(idx, wi) and torch.equal(score.view(torch.int32), wsc.view(torch.int32))):
                bad.append(t)
        ring.run(batches, rated, on_result=check)
        ring.wait()
        flag = torch.tensor([int(not bad)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["ring_equals_unsharded_bitwise_on_every_owner"] = bool(flag.item())
        ring.run([Ub[t % 4] for t in range(2 * world)])
        ring.wait(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ring.run([Ub[t % 4] for t in range(K)])
        ring.wait()
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out["ring"] = {"ms_per_step": float(ms.item()), "users_per_s": nb / (float(ms.item()) / 1e3), "steps": K}
        if rank == 0:
            print(json.dumps(out), flush=True)
        ring.close()
        del ring, V
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Row-sharded ALS over N GPUs (topkrec.dist.ShardedAls, NCCL): bit-identity with the single-GPU iteration + timing.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/als_2gpu.py [scale=0.125] [d=256]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "top-k-rec_b200"))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import topkrec  # noqa: E402
from topkrec import dist as tdist  # noqa: E402
from als_probe import synth  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_users, n_items = int(480189 * scale), 17770
    users, items = synth(n_users, n_items, 208)
    u_ptr = np.zeros(n_users + 1, np.int64); np.cumsum(np.bincount(users, minlength=n_users), out=u_ptr[1:])
    by_i = np.argsort(items, kind="stable")
    i_ptr = np.zeros(n_items + 1, np.int64); np.cumsum(np.bincount(items, minlength=n_items), out=i_ptr[1:])
    u_idx, i_idx = items.astype(np.int32), users[by_i].astype(np.int32)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    U0 = torch.rand(n_users, d, device="cuda", generator=g); V0 = torch.rand(n_items, d, device="cuda", generator=g)
    eng = tdist.ShardedAls(u_ptr, u_idx, i_ptr, i_idx, seg=4096)
    U, V = U0.clone(), V0.clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times, losses = [], None
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record()
        losses = eng.iteration(U, V, 1.0, 0.01, 0.01, 0.01, wmf=True)
        ev[1].record(); torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1])], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    # the same four iterations on this GPU alone
    us, its = topkrec.AlsSide(u_ptr, u_idx, 4096), topkrec.AlsSide(i_ptr, i_idx, 4096)
    U1, V1 = U0.clone(), V0.clone()
    t1 = []
    for it in range(4):
        torch.cuda.synchronize(); ev[0].record()
        l_u = topkrec.als_solve_rows(us, V1, U1, topkrec.als_gram(V1, its.rated_dev, 0.01, 0.01), 1.0, 0.01, 0.0, 0.01)
        l_i = topkrec.als_solve_rows(its, U1, V1, topkrec.als_gram(U1, us.rated_dev, 0.01, 0.0), 1.0, 0.01, 0.01, 0.01, item_loss=True)
        ev[1].record(); torch.cuda.synchronize(); t1.append(ev[0].elapsed_time(ev[1]))
    same = bool(torch.equal(U, U1) and torch.equal(V, V1))
    ok = torch.tensor([int(same)], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "users": n_users, "items": n_items, "d": d, "positives": int(users.size),
                          "bit_identical_to_one_gpu_on_every_rank": bool(int(ok)), "sharded_iteration_ms": times[1:], "one_gpu_iteration_ms": t1[1:],
                          "speedup": float(np.mean(t1[1:]) / np.mean(times[1:])), "user_bounds": eng.bounds[0], "item_bounds": eng.bounds[1],
                          "loss_terms_sharded": losses, "loss_terms_one_gpu": [float(l_u.sum()), float(l_i.sum())]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

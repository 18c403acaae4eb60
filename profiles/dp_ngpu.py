#!/usr/bin/env python3
"""Data-parallel K1 on N GPUs (torchrun): the fused peer-memory exchange (tkr_bpr_dp_step) against the NCCL all-reduce
route and against ONE GPU stepping the union batch, then device-timed steps of both routes on C2 (70k x 10k, d=128).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/dp_ngpu.py [batch] [timed_steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "top-k-rec_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import topkrec  # noqa: E402
from topkrec import dist as tdist  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    NU, NI, D = bench.N_USERS, bench.N_ITEMS, bench.D
    tr_users, indptr, pos_idx = bench.synth_interactions()
    smp = topkrec.Sampler(tdist.user_partition(tr_users, rank, world), indptr, pos_idx, NI, seed=123, device=dev)
    init = {k: torch.from_numpy(v).to(dev) for k, v in bench.init_state_np(NU, NI, D).items()}
    cfg = topkrec.BprCfg(NU, NI, D)
    hot = topkrec.popular_items(pos_idx, NI)
    steps = 4
    trip = [topkrec.bpr_sample(smp, (rank * 64 + t) * B, B, dev) for t in range(steps)]
    out = {"world": world, "batch_per_gpu": B}

    def run(exchange):
        st = {k: v.clone() for k, v in init.items()}
        eng = tdist.DataParallelBpr(cfg, st, B, exchange=exchange)
        topkrec.bpr_set_hot_items(cfg, B, eng.ws, hot)
        loss = torch.zeros(1, device=dev)
        for t in range(steps):
            eng.step(*trip[t], loss=loss)
        eng.check()
        eng.sync_slots()
        return eng, st, loss

    ef, sf, lf = run("peer")
    en, sn, ln = run("nccl")
    # 1. the two routes agree (same per-rank partial sums, another summation order across ranks)
    out["fused_vs_nccl_rel"] = {k: rel(sf[k], sn[k]) for k in ("V", "b", "U", "msV", "msb", "msU")}
    # 2. replicas bit-identical
    for k in ("V", "b"):
        hi, lo = sf[k].clone(), sf[k].clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        out["replicas_identical_" + k] = bool(torch.equal(hi, lo))
    # 3. one GPU stepping the union batch (rank 0 gathers every rank's triples)
    allt = [[torch.empty(world * B, dtype=torch.int32, device=dev) for _ in range(3)] for _ in range(steps)]
    for t in range(steps):
        for c in range(3):
            dist.all_gather_into_tensor(allt[t][c], trip[t][c])
    Uall = sf["U"].clone(); msUall = sf["msU"].clone()     # every rank's own user rows -> one table
    own = torch.zeros(NU, dtype=torch.bool, device=dev); own[rank::world] = True
    Uall[~own] = 0; msUall[~own] = 0
    dist.all_reduce(Uall); dist.all_reduce(msUall)
    if rank == 0:
        s1 = {k: v.clone() for k, v in init.items()}
        ws1 = topkrec.bpr_workspace(cfg, world * B, dev)
        topkrec.bpr_set_hot_items(cfg, world * B, ws1, hot)
        for t in range(steps):
            topkrec.bpr_step(cfg, s1["U"], s1["V"], s1["b"], s1["msU"], s1["msV"], s1["msb"], *allt[t], world * B, 1, ws1)
        out["fused_vs_one_gpu_union_batch_rel"] = {"V": rel(sf["V"], s1["V"]), "b": rel(sf["b"], s1["b"]), "msV": rel(sf["msV"], s1["msV"]),
                                                   "msb": rel(sf["msb"], s1["msb"]), "U": rel(Uall, s1["U"]), "msU": rel(msUall, s1["msU"])}
        del s1, ws1
    # 4. device-timed steps of both routes, sampler fused
    for name, eng in (("fused", ef), ("nccl", en)):
        for t in range(5):
            eng.step(sampler=smp, first_draw=(1 << 40) + t * B)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(K):
            eng.step(sampler=smp, first_draw=(2 << 40) + t * B)
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        eng.check()
        out[name + "_ms_per_step"] = float(ms.item())
        out[name + "_triples_per_s"] = world * B / (float(ms.item()) / 1e3)
    ef.close(); en.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in topkrec.dist: shard bounds,
the single-exchange candidate merge, and the data-parallel BPR decomposition.  The CUDA
kernels cannot run here, so the per-rank compute is injected from the oracle -- it is the
exchange / partition logic that is under test (the kernels themselves are covered by -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from topkrec import dist as tdist
from oracle import bpr_ref, topk_ref


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_score(U, V, k, bias, rp, ri, col_offset=0):
    n = lambda t: None if t is None else t.numpy()  # noqa: E731
    i, s = topk_ref.score_topk(U.numpy(), V.numpy(), k, n(bias), n(rp), n(ri), col_offset=col_offset)
    return torch.from_numpy(i), torch.from_numpy(s)


def _oracle_merge(idx, score):
    i, s = topk_ref.topk_merge(idx.numpy(), score.numpy())
    return torch.from_numpy(i), torch.from_numpy(s)


def _worker_topk(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                       # same data on both ranks
    U = torch.from_numpy(rng.standard_normal((23, 16)).astype(np.float32))
    V = rng.standard_normal((301, 16)).astype(np.float32); V[300] = V[0]; V[150] = V[149]
    bias = rng.standard_normal(301).astype(np.float32)
    rp = torch.from_numpy(np.arange(0, 24 * 5, 5, dtype=np.int64))
    ri = torch.from_numpy(np.sort(rng.integers(0, 301, (23, 5)), axis=1).astype(np.int32).ravel())
    beg, end = tdist.shard_bounds(301, world)[rank]
    idx, score = tdist.sharded_score_topk(U, torch.from_numpy(V[beg:end].copy()), 10, beg, torch.from_numpy(bias[beg:end].copy()),
                                          rp, ri, score_fn=_oracle_score, merge_fn=_oracle_merge)
    wi, ws = topk_ref.score_topk(U.numpy(), V, 10, bias, rp.numpy(), ri.numpy())
    q.put((rank, bool(np.array_equal(idx.numpy(), wi) and np.array_equal(score.numpy(), ws))))
    dist.destroy_process_group()


def _worker_dp(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    nu, ni, d, B = 40, 15, 8, 64
    st = bpr_ref.new_state(nu, ni, d, rng)
    whole = {k: v.copy() for k, v in st.items()}
    tr_users = np.arange(nu)
    mine = tdist.user_partition(tr_users, rank, world)
    assert (mine % world == rank).all() and mine.size == nu // world
    # every rank draws the SAME global batch and keeps the triples of its own users
    u = rng.integers(0, nu, world * B); i = rng.integers(0, ni, world * B); j = rng.integers(0, ni, world * B)
    keep = u % world == rank
    cfg = bpr_ref.BprCfg()
    _, _, gU, gVi, gVj, gbi, gbj = bpr_ref.bpr_occurrence_grads(st["U"], st["V"], st["b"], u[keep], i[keep], j[keep], cfg)
    GV = np.zeros((ni, d), np.float32); Gb = np.zeros(ni, np.float32); tch = np.zeros(ni, np.float32)
    np.add.at(GV, i[keep], gVi); np.add.at(GV, j[keep], gVj); np.add.at(Gb, i[keep], gbi); np.add.at(Gb, j[keep], gbj)
    np.add.at(tch, i[keep], 1); np.add.at(tch, j[keep], 1)
    flat = torch.from_numpy(np.concatenate([GV.ravel(), Gb, tch]))          # the contiguous [GV | Gb | tchV] region
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat = flat.numpy()
    GV, Gb, tch = flat[:ni * d].reshape(ni, d), flat[ni * d:ni * d + ni], flat[ni * d + ni:]
    rows = np.nonzero(tch)[0]
    bpr_ref.apply_sparse(st["V"], st["msV"], rows, GV[rows], cfg)
    bpr_ref.apply_sparse(st["b"], st["msb"], rows, Gb[rows], cfg)
    rU, GU = bpr_ref.segment_sum(u[keep], gU)
    bpr_ref.apply_sparse(st["U"], st["msU"], rU, GU, cfg)
    bpr_ref.bpr_step(whole, u, i, j, cfg)                                    # one GPU, batch world*B
    ok = all(np.abs(st[n] - whole[n]).max() <= 1e-6 * np.abs(whole[n]).max() for n in ("V", "b", "msV", "msb"))
    own = np.arange(nu) % world == rank
    ok = ok and all(np.abs(st[n][own] - whole[n][own]).max() <= 1e-6 * np.abs(whole[n]).max() for n in ("U", "msU"))
    # replicas must hold identical item tables after the step
    t = torch.from_numpy(st["V"].copy()); ref = t.clone(); dist.broadcast(ref, 0)
    q.put((rank, bool(ok and torch.equal(t, ref))))
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker_topk, _worker_dp])
def test_world2_gloo(worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == [(0, True), (1, True)]


def test_shard_bounds():
    assert tdist.shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert tdist.shard_bounds(1 << 20, 8)[-1] == (7 << 17, 1 << 20)
    b = tdist.shard_bounds(5, 8)
    assert b[0] == (0, 1) and b[-1] == (5, 5) and sum(e - s for s, e in b) == 5

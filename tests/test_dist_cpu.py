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
from oracle import als_ref, bpr_ref, topk_ref


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


def _worker_dp_owner(rank, world, port, q):
    """the fused exchange of tkr_bpr_dp_step, emulated: every rank keeps its partial [GV|Gb|tch]; the OWNER of an item row
    (row % world) reads that row from every rank in rank order, sums, applies the optimiser once with its local slot, and
    writes the new row into every replica.  Must equal one process stepping the union batch, with identical replicas."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(6)
    nu, ni, d, B = 30, 13, 8, 48
    st = bpr_ref.new_state(nu, ni, d, rng)
    st["b"] = (0.01 * rng.standard_normal(ni)).astype(np.float32)
    whole = {k: v.copy() for k, v in st.items()}
    cfg = bpr_ref.BprCfg(lambda_b=0.01)
    ok = True
    for step in range(3):
        u = rng.integers(0, nu, world * B); i = rng.integers(0, ni - 2, world * B); j = rng.integers(0, ni - 2, world * B)   # two items never touched
        keep = u % world == rank
        _, _, gU, gVi, gVj, gbi, gbj = bpr_ref.bpr_occurrence_grads(st["U"], st["V"], st["b"], u[keep], i[keep], j[keep], cfg)
        G = np.zeros((ni, d + 2), np.float32)                              # [GV | Gb | tch] per item row
        np.add.at(G[:, :d], i[keep], gVi); np.add.at(G[:, :d], j[keep], gVj)
        np.add.at(G[:, d], i[keep], gbi); np.add.at(G[:, d], j[keep], gbj)
        np.add.at(G[:, d + 1], i[keep], 1); np.add.at(G[:, d + 1], j[keep], 1)
        parts = [torch.zeros(ni, d + 2) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(G))                        # stands in for the NVLink loads of the peers' rows
        newV, newb = np.zeros_like(st["V"]), np.zeros_like(st["b"])
        for r in range(rank, ni, world):                                   # owned rows
            tot = np.zeros(d + 2, np.float32)
            for p in range(world):
                tot += parts[p][r].numpy()
            if tot[d + 1] == 0:
                newV[r], newb[r] = st["V"][r], st["b"][r]
                continue
            rows = np.array([r])
            bpr_ref.apply_sparse(st["V"], st["msV"], rows, tot[None, :d], cfg)
            bpr_ref.apply_sparse(st["b"], st["msb"], rows, tot[d:d + 1], cfg)
            newV[r], newb[r] = st["V"][r], st["b"][r]
        tv, tb = torch.from_numpy(newV), torch.from_numpy(newb)
        dist.all_reduce(tv); dist.all_reduce(tb)                           # stands in for the NVLink stores into every replica (disjoint rows)
        st["V"][:], st["b"][:] = tv.numpy(), tb.numpy()
        rU, GU = bpr_ref.segment_sum(u[keep], gU)
        bpr_ref.apply_sparse(st["U"], st["msU"], rU, GU, cfg)
        bpr_ref.bpr_step(whole, u, i, j, cfg)
        ok = ok and all(np.abs(st[n] - whole[n]).max() <= 2e-6 * np.abs(whole[n]).max() for n in ("V", "b"))
        own_items = np.arange(ni) % world == rank
        ok = ok and all(np.abs(st[n][own_items] - whole[n][own_items]).max() <= 2e-6 * np.abs(whole[n]).max() for n in ("msV", "msb"))
        own = np.arange(nu) % world == rank
        ok = ok and all(np.abs(st[n][own] - whole[n][own]).max() <= 2e-6 * np.abs(whole[n]).max() for n in ("U", "msU"))
        t = torch.from_numpy(st["V"].copy()); ref = t.clone(); dist.broadcast(ref, 0)
        ok = ok and torch.equal(t, ref)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_slices(rank, world, port, q):
    """user-slice ownership of the item-sharded scorer (tkr_topk_exchange_*): owner of row u = u // ceil(n / world); every
    rank's shard lists of a slice, merged by the owner, equal the unsharded lists; the slices tile the batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    n, ni, d, k = 37, 203, 12, 10
    U = rng.standard_normal((n, d)).astype(np.float32); V = rng.standard_normal((ni, d)).astype(np.float32)
    beg, end = tdist.shard_bounds(ni, world)[rank]
    li, ls = topk_ref.score_topk(U, V[beg:end], k, col_offset=beg)          # this rank's lists of ALL users
    sl = -(-n // world)
    lo, hi = min(n, rank * sl), min(n, (rank + 1) * sl)
    gi = [torch.zeros((n, k), dtype=torch.int32) for _ in range(world)]; gs = [torch.zeros((n, k)) for _ in range(world)]
    dist.all_gather(gi, torch.from_numpy(li)); dist.all_gather(gs, torch.from_numpy(ls))
    mi, ms = topk_ref.topk_merge(np.stack([g[lo:hi].numpy() for g in gi]), np.stack([g[lo:hi].numpy() for g in gs]))
    wi, ws = topk_ref.score_topk(U[lo:hi], V, k)
    cover = torch.zeros(n); cover[lo:hi] = 1
    dist.all_reduce(cover)
    q.put((rank, bool(np.array_equal(mi, wi) and np.array_equal(ms, ws) and bool((cover == 1).all()))))
    dist.destroy_process_group()


def _worker_ring(rank, world, port, q):
    """the ring of sweep segments (topkrec.dist.RingScorer) on gloo: the running state of a batch's sweep -- here its running
    top-k lists -- travels rank to rank in the slot order of ``ring_slot``; rank r always sweeps item shard r; the rank that
    holds the last segment (``ring_owner``) ends with the unsharded lists.  Per-segment compute injected from the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    n, ni, d, k, T = 19, 157, 8, 10, 5
    V = rng.standard_normal((ni, d)).astype(np.float32); V[100] = V[3]
    batches = [rng.standard_normal((n, d)).astype(np.float32) for _ in range(T)]
    beg, end = tdist.shard_bounds(ni, world)[rank]
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    ok, owned, pending = True, [], []
    for tau in range(T + world - 1):
        slot = tdist.ring_slot(tau, rank, world, T)
        if slot is None:
            continue
        t, p = slot
        ok = ok and (2 * t + p) % world == rank                         # batch t started at rank 2t mod G, one rank per slot
        li, ls = topk_ref.score_topk(batches[t], V[beg:end], k, col_offset=beg)
        if p > 0:
            st = torch.empty((2, n, k), dtype=torch.int32)
            dist.recv(st, src=prev, tag=t)
            li, ls = topk_ref.topk_merge(np.stack([st[0].numpy(), li]), np.stack([st[1].view(torch.float32).numpy(), ls]))
        if p < world - 1:
            pending.append(dist.isend(torch.stack([torch.from_numpy(li), torch.from_numpy(ls).view(torch.int32)]), dst=nxt, tag=t))
        else:
            wi, ws = topk_ref.score_topk(batches[t], V, k)
            ok = ok and tdist.ring_owner(t, world) == rank and np.array_equal(li, wi) and np.array_equal(ls, ws)
            owned.append(t)
    for h in pending:
        h.wait()
    seen = torch.zeros(T); seen[owned] = 1
    dist.all_reduce(seen)
    q.put((rank, bool(ok and bool((seen == 1).all()))))
    dist.destroy_process_group()


class _CpuSide:
    def __init__(self, indptr, idx):
        self.indptr, self.idx = indptr, idx


def _oracle_gram(Y, rows, scale, ridge):
    return torch.from_numpy(als_ref.shared_gram(Y.numpy(), rows.numpy(), scale, ridge))


def _oracle_solve(side, Y, X, base, a, b, ridge, lreg, prior=None, solve_empty=False, item_loss=False):
    """per-row restatement with the matrix handed in (cer.py:39-45 / :49-62) on a CPU slice, in place"""
    Yn, Xn, Bn = Y.numpy(), X.numpy(), base.numpy()
    k = Yn.shape[1]
    loss = np.zeros(Xn.shape[0])
    for r in range(Xn.shape[0]):
        pos = side.idx[side.indptr[r]:side.indptr[r + 1]]
        if len(pos) or solve_empty:
            Yi = Yn[pos]
            rhs = np.sum(Yi, axis=0) * a + (prior[r].numpy() * ridge if prior is not None else 0)
            Xn[r] = np.linalg.solve(np.dot(Yi.T, Yi) * (a - b) + Bn + np.eye(k, dtype=np.float32) * ridge, rhs)
        loss[r] = 0.5 * lreg * np.sum(Xn[r] ** 2)
    return torch.from_numpy(loss)


def _worker_als(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    nu, ni, d = 37, 21, 6
    cnt = rng.integers(0, 9, nu); cnt[3] = 40
    u_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    u_idx = rng.integers(0, ni - 2, int(u_ptr[-1])).astype(np.int32)
    users = np.repeat(np.arange(nu), cnt)
    by_i = np.argsort(u_idx, kind="stable")
    i_ptr = np.zeros(ni + 1, np.int64); np.cumsum(np.bincount(u_idx, minlength=ni), out=i_ptr[1:])
    i_idx = users[by_i].astype(np.int32)
    U0, V0 = rng.random((nu, d)).astype(np.float32), rng.random((ni, d)).astype(np.float32)
    eng = tdist.ShardedAls(u_ptr, u_idx, i_ptr, i_idx, side_fn=_CpuSide, gram_fn=_oracle_gram, solve_fn=_oracle_solve)
    bounds = eng.bounds[0]
    ok = bounds[0][0] == 0 and bounds[-1][1] == nu and all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    U, V = torch.from_numpy(U0.copy()), torch.from_numpy(V0.copy())
    for _ in range(2):
        eng.iteration(U, V, 1.0, 0.01, 0.01, 0.01, wmf=True)
    # unsharded: the oracle's own half-steps
    ru, rv = U0.copy(), V0.copy()
    u_rated = np.flatnonzero(np.diff(u_ptr) > 0); i_rated = np.flatnonzero(np.diff(i_ptr) > 0)
    for _ in range(2):
        als_ref.user_step(ru, rv, u_ptr, u_idx, i_rated, 1.0, 0.01, 0.01)
        als_ref.item_step(ru, rv, i_ptr, i_idx, u_rated, 1.0, 0.01, 0.01, None)
    ok = ok and np.array_equal(U.numpy(), ru) and np.array_equal(V.numpy(), rv)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker_topk, _worker_dp, _worker_dp_owner, _worker_slices, _worker_ring, _worker_als])
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


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_ring_schedule(G):
    """every slot holds at most one batch per rank; a batch visits G consecutive ranks in G consecutive slots, segment p at
    rank (2t + p) mod G; owners rotate over the ring"""
    T = 2 * G + 3
    visits = {}
    for tau in range(T + G - 1):
        here = [tdist.ring_slot(tau, r, G, T) for r in range(G)]
        ts = [s[0] for s in here if s is not None]
        assert len(ts) == len(set(ts))
        for r, s in enumerate(here):
            if s is not None:
                visits.setdefault(s[0], []).append((tau, r, s[1]))
    assert sorted(visits) == list(range(T))
    for t, v in visits.items():
        assert [x[0] for x in v] == list(range(t, t + G)) and [x[2] for x in v] == list(range(G))
        assert [x[1] for x in v] == [(2 * t + p) % G for p in range(G)] and v[-1][1] == tdist.ring_owner(t, G)
    if G > 1 and G % 2 == 0:
        assert {tdist.ring_owner(t, G) for t in range(T)} == set(range(1, G, 2)) or G == 2
    if G % 2 == 1:
        assert {tdist.ring_owner(t, G) for t in range(T)} == set(range(G))


def test_ring_shard_bounds():
    assert tdist.ring_shard_bounds(1 << 20, 8) == tdist.shard_bounds(1 << 20, 8)          # equal shards by default (measured best)
    assert tdist.ring_shard_bounds(1 << 20, 3, 0.05) == tdist.shard_bounds(1 << 20, 3)    # odd rings rotate every role: no skew
    for G, tail in ((2, 0.03), (4, 0.03), (8, 0.04), (8, -0.02)):
        b = tdist.ring_shard_bounds(1 << 20, G, tail)
        assert b[0][0] == 0 and b[-1][1] == 1 << 20 and all(b[r][1] == b[r + 1][0] for r in range(G - 1))
        assert all(e % 256 == 0 for _, e in b)
        even, odd = b[0][1] - b[0][0], b[1][1] - b[1][0]
        assert (even > odd) == (tail > 0)


def test_shard_bounds():
    assert tdist.shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert tdist.shard_bounds(1 << 20, 8)[-1] == (7 << 17, 1 << 20)
    b = tdist.shard_bounds(5, 8)
    assert b[0] == (0, 1) and b[-1] == (5, 5) and sum(e - s for s, e in b) == 5


def test_balanced_row_bounds():
    indptr = np.concatenate([[0], np.cumsum([1000] + [10] * 99)])
    b = tdist.balanced_row_bounds(indptr, 4, row_cost=0)
    assert b[0][0] == 0 and b[-1][1] == 100 and all(b[r][1] == b[r + 1][0] for r in range(3))
    work = [indptr[e] - indptr[s] for s, e in b]
    assert max(work) <= 1000 + 10 and sum(work) == indptr[-1]            # the heavy row sits alone
    assert tdist.balanced_row_bounds(np.array([0, 0, 0]), 4) [-1][1] == 2

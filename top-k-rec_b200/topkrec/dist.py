"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* score + top-k: item columns are sharded contiguously; every rank scores the same
  user batch against its shard (columns reported globally via ``col_offset``), the
  per-shard candidate lists are exchanged with ONE all-gather and merged with the
  (score desc, column desc) order, so the result is bit-identical to one GPU.
* BPR: users are partitioned ``u % world == rank`` (U rows and their slots never
  move); V/b are replicated; per step the contiguous fp32 region [GV | Gb | tchV]
  of the workspace is all-reduced between ``tkr_bpr_grad`` and ``tkr_bpr_apply``.
  Equivalent to a single-GPU batch of world*B triples up to fp32 summation order.

The reference is single-process; these semantics are defined by
"same answer as the unsharded call".
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, balanced [beg, end) per rank; the first n % world shards get one extra."""
    base, extra = divmod(int(n), int(world))
    bounds, beg = [], 0
    for r in range(world):
        end = beg + base + (1 if r < extra else 0)
        bounds.append((beg, end))
        beg = end
    return bounds


def user_partition(tr_users, rank, world):
    """Users a rank samples from / owns the rows of."""
    tr_users = np.asarray(tr_users)
    return tr_users[tr_users % world == rank]


def _gather(t, group):
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def sharded_score_topk(U, V_shard, k, col_offset, bias_shard=None, rated_indptr=None, rated_idx=None, group=None,
                       score_fn=None, merge_fn=None, engine="tc", ws=None, items_prepared=False):
    """Filtered top-k over item-sharded V.  ``score_fn`` / ``merge_fn`` default to the CUDA engine
    (tests on CPU/gloo inject checkers to exercise the exchange logic)."""
    if score_fn is None or merge_fn is None:
        import topkrec
        score_fn = score_fn or (lambda *a, **kw: topkrec.score_topk(*a, engine=engine, ws=ws, items_prepared=items_prepared, **kw))
        merge_fn = merge_fn or topkrec.topk_merge
    idx, score = score_fn(U, V_shard, k, bias_shard, rated_indptr, rated_idx, col_offset=col_offset)
    if group is None and not dist.is_initialized():
        return idx, score
    # one exchange: pack (idx, score bits) as int32 [2, nu, k]
    packed = torch.stack([idx, score.view(torch.int32)])
    allp = _gather(packed, group)
    return merge_fn(allp[:, 0].contiguous(), allp[:, 1].contiguous().view(torch.float32))


class DataParallelBpr:
    """Synchronous data-parallel BPR step over the ranks of ``group``."""

    def __init__(self, cfg, state, batch, group=None):
        import topkrec
        self.t = topkrec
        self.cfg, self.st, self.batch, self.group = cfg, state, int(batch), group
        self.ws = topkrec.bpr_workspace(cfg, batch, state["U"].device)
        self.item_grads = topkrec.bpr_item_grad_view(cfg, batch, self.ws)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def step(self, u=None, i=None, j=None, sampler=None, first_draw=0, loss=None):
        st, t = self.st, self.t
        dp = self.world > 1
        t.bpr_grad(self.cfg, st["U"], st["V"], st["b"], u, i, j, self.batch, self.ws, loss, sampler, first_draw, data_parallel=dp)
        if dp:
            dist.all_reduce(self.item_grads, op=dist.ReduceOp.SUM, group=self.group)
        t.bpr_apply(self.cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], self.batch, self.ws, data_parallel=dp)


def balanced_row_bounds(indptr, world, row_cost=256):
    """Contiguous [beg, end) row ranges per rank with about equal work, a row costing its positives + ``row_cost``
    (the factorisation is a fixed cost per row; SURVEY.md 8(e): U-step sharded by user, V-step by item)."""
    indptr = np.asarray(indptr, np.int64)
    n = indptr.size - 1
    w = np.cumsum(np.diff(indptr) + row_cost)
    bounds, beg, before = [], 0, 0
    for r in range(world):
        if r == world - 1 or beg >= n:
            end = n
        else:
            target = (w[-1] - before) / (world - r)
            end = min(n, int(np.searchsorted(w, before + target, side="left")) + 1)      # up to and including the crossing row
            if end - 1 > beg and (before + target) - w[end - 2] < w[end - 1] - (before + target):
                end -= 1                                                                 # ... unless leaving it out is closer
        bounds.append((beg, end))
        before = w[end - 1] if end > 0 else 0
        beg = end
    return bounds


class ShardedAls:
    """One ALS iteration (single/cer.py:36-63 / the intended single/wmf.py:67-96) over the ranks of ``group``: every rank
    holds both factors, solves its contiguous block of user rows, the blocks are exchanged (one broadcast per rank, in
    place), then the same for the items.  The shared Gram is computed by every rank from the full factor, so all
    replicas stay bit-identical and equal the single-GPU result (rows are independent given the other factor).

    ``side_fn(indptr, idx)`` / ``gram_fn`` / ``solve_fn`` default to the CUDA engine; the gloo tests inject the oracle."""

    def __init__(self, u_ptr, u_idx, i_ptr, i_idx, group=None, seg=1024, device="cuda", side_fn=None, gram_fn=None, solve_fn=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if side_fn is None:
            import topkrec
            side_fn = lambda p, i: topkrec.AlsSide(p, i, seg, device)  # noqa: E731
            gram_fn, solve_fn = topkrec.als_gram, topkrec.als_solve_rows
        self.gram_fn, self.solve_fn = gram_fn, solve_fn
        self.sides, self.bounds, self.rated = [], [], []
        for ptr, idx in ((np.asarray(u_ptr, np.int64), np.asarray(u_idx, np.int32)), (np.asarray(i_ptr, np.int64), np.asarray(i_idx, np.int32))):
            b = balanced_row_bounds(ptr, self.world)
            beg, end = b[self.rank]
            self.bounds.append(b)
            self.sides.append(side_fn(ptr[beg:end + 1] - ptr[beg], idx[ptr[beg]:ptr[end]]))
            self.rated.append(np.flatnonzero(np.diff(ptr) > 0).astype(np.int32))
        self._rated_dev = None

    def _exchange(self, X, bounds):
        if self.world > 1:
            for r, (beg, end) in enumerate(bounds):
                if end > beg:
                    dist.broadcast(X[beg:end], src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)

    def iteration(self, U, V, a, b, lu, lv, prior=None, wmf=False):
        """Updates U and V in place on every rank; returns the loss terms of the two half-steps (summed over ranks)."""
        if self._rated_dev is None:
            self._rated_dev = [torch.from_numpy(r).to(U.device) for r in self.rated]
        (ub, ue), (ib, ie) = self.bounds[0][self.rank], self.bounds[1][self.rank]
        XX = self.gram_fn(V, self._rated_dev[1], b, lu)
        l_u = self.solve_fn(self.sides[0], V, U[ub:ue], XX, a, b, 0.0, lu).sum()
        self._exchange(U, self.bounds[0])
        XXv = self.gram_fn(U, self._rated_dev[0], b, 0.0)
        l_i = self.solve_fn(self.sides[1], U, V[ib:ie], XXv, a, b, lv, lv, prior=None if prior is None else prior[ib:ie],
                            solve_empty=prior is not None and not wmf, item_loss=True).sum()
        self._exchange(V, self.bounds[1])
        loss = torch.stack([l_u, l_i]).double()
        if self.world > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return float(loss[0]), float(loss[1])

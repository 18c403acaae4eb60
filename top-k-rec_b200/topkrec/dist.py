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

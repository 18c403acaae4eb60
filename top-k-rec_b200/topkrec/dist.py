"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* score + top-k: item columns are sharded contiguously; every rank scores the same
  user batch against its shard (columns reported globally via ``col_offset``), the
  per-shard candidate lists are exchanged with ONE all-gather and merged with the
  (score desc, column desc) order, so the result is bit-identical to one GPU.
* BPR: users are partitioned ``u % world == rank`` (U rows and their slots never
  move); V/b are replicated; per step the item gradients are exchanged and applied by
  ONE kernel over peer memory (``tkr_bpr_dp_step``: NVLink loads of the owned rows of
  every rank's accumulator, one update, NVLink stores of the new rows into every replica).
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


class ShardedScorer:
    """Item-sharded filtered top-k with the candidate exchange on peer memory (``tkr_topk_exchange_*``).

    Every rank scores the whole user batch against its item shard on the main stream, pushes each user's local list
    into the merge buffer of the user's owner (rank = row // ceil(n / world)) with NVLink stores, and merges the lists
    of its own slice on a side stream -- so the exchange + merge of batch t overlap the scoring of batch t+1.  The
    final list of a user lives on exactly one rank (``submit`` returns this rank's slice) and equals the unsharded
    result bit for bit.  All ranks must call ``submit`` with batches of the same number of rows, in the same order."""

    def __init__(self, n_items_shard, d, k, user_batch, col_offset, group=None, engine="tc", has_bias=False, device=None):
        import topkrec
        from .peer import PeerBuffer
        self.t, self.group = topkrec, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.ni, self.d, self.k, self.nb, self.engine, self.col_offset = int(n_items_shard), int(d), int(k), int(user_batch), engine, int(col_offset)
        L = topkrec.lib()
        need = (L.tkr_score_topk_tc_workspace_bytes(self.nb, self.ni, self.d, self.k, int(has_bias)) if engine == "tc"
                else L.tkr_score_topk_workspace_bytes(self.nb, self.ni, self.d, self.k))
        with torch.cuda.device(self.dev):
            self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.dev)
            self.buf = PeerBuffer(L.tkr_topk_exchange_bytes(self.nb, self.k, self.world), group, self.dev)
            self.side = torch.cuda.Stream(self.dev)
            self.slice = -(-self.nb // self.world)
            self.local = [(torch.empty((self.nb, k), dtype=torch.int32, device=self.dev), torch.empty((self.nb, k), dtype=torch.float32, device=self.dev)) for _ in range(2)]
            self.out = [(torch.empty((self.slice, k), dtype=torch.int32, device=self.dev), torch.empty((self.slice, k), dtype=torch.float32, device=self.dev)) for _ in range(2)]
            self.pushed = [torch.cuda.Event() for _ in range(2)]
            self.merged = [torch.cuda.Event() for _ in range(2)]
        self.epoch = 0
        self.prepared = False
        if self.world > 1:
            dist.barrier(group=group)

    def rows_of(self, n, rank=None):
        """[beg, end) of the user rows of a batch of n whose final lists live on ``rank`` (default: this rank)"""
        rank = self.rank if rank is None else rank
        sl = -(-n // self.world)
        return min(n, rank * sl), min(n, (rank + 1) * sl)

    def submit(self, U, V_shard, bias_shard=None, rated_indptr=None, rated_idx=None):
        """Enqueue one user batch; returns (idx, score) device views of this rank's slice, valid on the main stream
        after ``self.merged[slot]`` (already made a dependency of the main stream before the slot is reused) --
        call ``wait()`` or read them on ``self.side``.  ``V_shard`` must stay the same table between calls while
        ``self.prepared`` (its BF16 copy is reused)."""
        t, L = self.t, self.t.lib()
        n = U.shape[0]
        if n > self.nb:
            raise ValueError("batch of %d rows exceeds the scorer's user_batch %d" % (n, self.nb))
        self.epoch += 1
        s = self.epoch & 1
        main = torch.cuda.current_stream(self.dev)
        if self.epoch > 2:
            main.wait_event(self.merged[s])            # the merge of epoch-2 has consumed local[s] / out[s] is free again
        li, ls = self.local[s][0][:n], self.local[s][1][:n]
        t.score_topk(U, V_shard, self.k, bias_shard, rated_indptr, rated_idx, col_offset=self.col_offset, engine=self.engine, ws=self.ws,
                     out=(li, ls), items_prepared=self.prepared and self.engine == "tc" and n == self.nb)
        self.prepared = n == self.nb
        with torch.cuda.device(self.dev):
            t._lib._check(L.tkr_topk_exchange_push(li.data_ptr(), ls.data_ptr(), n, self.nb, self.k, self.buf.peers_ptr, self.epoch, main.cuda_stream))
            self.pushed[s].record(main)
            beg, end = self.rows_of(n)
            oi, osc = self.out[s][0][:end - beg], self.out[s][1][:end - beg]
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.pushed[s])
                t._lib._check(L.tkr_topk_exchange_merge(n, self.nb, self.k, self.buf.peers_ptr, self.epoch, oi.data_ptr(), osc.data_ptr(), self.side.cuda_stream))
                self.merged[s].record(self.side)
        return oi, osc

    def wait(self):
        """main stream waits for every merge submitted so far; raises if a cross-GPU barrier timed out"""
        main = torch.cuda.current_stream(self.dev)
        main.wait_stream(self.side)
        with torch.cuda.device(self.dev):
            self.t._lib._check(self.t.lib().tkr_topk_exchange_status(self.nb, self.k, self.buf.peers_ptr, main.cuda_stream))

    def close(self):
        self.buf.close()


def ring_slot(tau, rank, world, n_batches):
    """Schedule of the ring of sweep segments: the batch rank ``rank`` works on in time slot ``tau`` and which segment of that
    batch's sweep it is.  Batch t starts at rank ``2 t mod world`` in slot t and moves one rank per slot, so in slot tau it sits
    at rank ``(tau + t) mod world`` -- distinct for the ``world`` batches in flight.  Returns ``(t, p)`` with p in
    [0, world) the segment number (0 = first, world-1 = last), or ``None`` for an idle slot (pipeline fill / drain)."""
    t = tau - ((2 * tau - rank) % world)
    if t < 0 or t >= n_batches:
        return None
    return t, tau - t


def ring_shard_bounds(n, world, tail_cost=0.0):
    """Item shards for ``RingScorer``: equal by default.  With an even number of ranks the sweeps START on the even ranks
    (threshold warm-up) and END on the odd ones (merge + exact re-scoring), so the two groups could be given shards of different
    lengths (odd ranks shorter by the fraction delta = tail_cost * world / 2, even ranks longer; negative: the other way round) --
    measured at 8 GPUs (profiles/r02z_ring_skew_8gpu_*.json): 0.712 ms per batch with equal shards, 0.749 / 0.765 / 0.785 at
    tail_cost 0.02 / 0.03 / 0.04 and 0.716 / 0.727 / 0.738 at -0.01 / -0.02 / -0.03: a slot lasts as long as its LONGEST shard
    on either side, neither role carries a visible fixed cost.  Kept for other shapes (few users per batch make the re-scoring
    relatively dearer).  Bounds are multiples of 256 items (the filter's tile width)."""
    world = int(world)
    if world < 2 or world % 2 == 1 or tail_cost == 0:
        return shard_bounds(n, world)
    delta = max(-0.5, min(0.5, 0.5 * tail_cost * world))
    w = [1.0 + delta if r % 2 == 0 else 1.0 - delta for r in range(world)]
    tot, acc, bounds, beg = sum(w), 0.0, [], 0
    for r in range(world):
        acc += w[r]
        end = int(n) if r == world - 1 else min(int(n), int(round(n * acc / tot / 256.0)) * 256)
        end = max(end, beg)
        bounds.append((beg, end))
        beg = end
    return bounds


def ring_owner(t, world):
    """rank that sweeps the last segment of batch t and ends up with its lists"""
    return (2 * t + world - 1) % world


class RingScorer:
    """Item-sharded filtered top-k as a RING of sweep segments (``tkr_score_topk_tc_segment``).

    The sweep of a user batch over the whole item table is cut into one segment per GPU (rank r always sweeps item shard r);
    the running state of the sweep (every row's threshold, candidate buffer and count, ~1 KB per row) travels from rank to
    rank over NVLink.  Unlike independent per-shard lists (``ShardedScorer``), the per-row selection work of a sweep -- which
    hardly depends on its length -- is then paid once per batch instead of once per shard.  Batch t STARTS at rank
    ``2 t mod G`` (the start of a sweep carries the threshold warm-up, the end the exact re-scoring: rotating them spreads
    both over the ring) and moves one rank per time slot, so in slot tau it is at rank ``(tau + t) mod G`` -- distinct for
    the G batches in flight.  The rank that holds a batch's last segment sorts its lists out, re-scores them exactly against
    the whole table (every rank keeps ``V_full`` for that) and owns the result: bit-identical to one GPU.
    All ranks call ``run`` with the same batches."""

    SLOT_READY, SLOT_FREE = 4, 5

    def __init__(self, V_shard, d, k, user_batch, col_offset, V_full, bias_shard=None, bias_full=None, group=None, device=None):
        import topkrec
        from .peer import PeerBuffer, _RawCuda
        self.t, self.group = topkrec, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.V, self.V_full, self.bias, self.bias_full = V_shard, V_full, bias_shard, bias_full
        self.d, self.k, self.nb, self.col_offset = int(d), int(k), int(user_batch), int(col_offset)
        L = topkrec.lib()
        self.sbytes = -(-L.tkr_score_topk_tc_state_bytes(self.nb) // 1024) * 1024
        self.flag_off = 2 * self.sbytes
        with torch.cuda.device(self.dev):
            self.buf = PeerBuffer(self.flag_off + L.tkr_peer_flag_bytes(), group, self.dev)
            self.ws = torch.empty(L.tkr_score_topk_tc_segment_workspace_bytes(self.nb, V_shard.shape[0], V_full.shape[0], self.d, self.k, int(bias_shard is not None)),
                                  dtype=torch.uint8, device=self.dev)
            self.xfer = torch.cuda.Stream(self.dev)
            self.seg_done = [torch.cuda.Event() for _ in range(2)]
            self.slot_free = [torch.cuda.Event() for _ in range(2)]
            self.out = [(torch.empty((self.nb, k), dtype=torch.int32, device=self.dev), torch.empty((self.nb, k), dtype=torch.float32, device=self.dev))
                        for _ in range(2)]
            self.nfb = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.state = [self.buf.local[s * self.sbytes:(s + 1) * self.sbytes] for s in range(2)]
            nxt = self.buf.ptrs[(self.rank + 1) % self.world]
            self.next_state = [torch.as_tensor(_RawCuda(nxt + s * self.sbytes, self.sbytes), device=self.dev) for s in range(2)] if self.world > 1 else None
        self.base = 0                 # time slots consumed by earlier run() calls (the flag epochs keep growing)
        self.prepared = False
        if self.world > 1:
            dist.barrier(group=group)

    def owner(self, t):
        """rank that ends up with the lists of batch t"""
        return ring_owner(t, self.world)

    def run(self, batches, rated=None, on_result=None):
        """``batches``: sequence of user batches (device tensors [user_batch, d], the same on every rank); ``rated``: optional
        sequence of (rated_indptr, rated_idx) per batch.  ``on_result(t, idx, score)`` is called (on the current stream, in
        stream order) on the rank that owns batch t; the views are reused two slots later."""
        t_, L, G, r = self.t, self.t.lib(), self.world, self.rank
        T = len(batches)
        main = torch.cuda.current_stream(self.dev)
        peers = self.buf.peers_ptr
        prev, nxt = (r - 1) % G, (r + 1) % G
        with torch.cuda.device(self.dev):
            for tau in range(T + G - 1):
                e, s = self.base + tau + 1, (self.base + tau) & 1
                slot = ring_slot(tau, r, G, T)
                if slot is None:
                    # idle slot (pipeline fill / drain): only the predecessor's "your slot is free" is owed, in slot order
                    if G > 1:
                        t_._lib._check(L.tkr_peer_signal_to(peers, self.flag_off, self.SLOT_FREE, prev, e, self.xfer.cuda_stream))
                    continue
                t, p = slot
                first, last = p == 0, p == G - 1
                if first:
                    main.wait_event(self.slot_free[s])                      # local slot s: its use two slots ago is over
                else:
                    t_._lib._check(L.tkr_peer_wait_from(peers, self.flag_off, self.SLOT_READY, prev, e, main.cuda_stream))
                rp, ri = rated[t] if rated is not None else (None, None)
                out = self.out[s] if last else None
                t_.score_topk_segment(batches[t], self.V, self.k, self.col_offset, self.state[s], first, last, V_full=self.V_full, bias_shard=self.bias,
                                      bias_full=self.bias_full, rated_indptr=rp, rated_idx=ri, out=out, ws=self.ws, n_fallback=self.nfb if last else None,
                                      items_prepared=self.prepared)
                self.prepared = True
                if last and on_result is not None:
                    on_result(t, out[0], out[1])
                self.seg_done[s].record(main)
                # everything that crosses to a neighbour goes through the transfer stream, in slot order: the "free" of slot tau is
                # only raised once every earlier slot's state has left this rank
                with torch.cuda.stream(self.xfer):
                    self.xfer.wait_event(self.seg_done[s])
                    if not last:
                        # the state lands in the next rank's buffer of time slot tau + 1: free once that rank is done with slot tau - 1
                        if self.base + tau >= 1:
                            t_._lib._check(L.tkr_peer_wait_from(peers, self.flag_off, self.SLOT_FREE, nxt, e - 1, self.xfer.cuda_stream))
                        self.next_state[(self.base + tau + 1) & 1].copy_(self.state[s], non_blocking=True)
                        t_._lib._check(L.tkr_peer_signal_to(peers, self.flag_off, self.SLOT_READY, nxt, e + 1, self.xfer.cuda_stream))
                    self.slot_free[s].record(self.xfer)
                    if G > 1:
                        t_._lib._check(L.tkr_peer_signal_to(peers, self.flag_off, self.SLOT_FREE, prev, e, self.xfer.cuda_stream))
        self.base += T + G - 1

    def wait(self):
        main = torch.cuda.current_stream(self.dev)
        main.wait_stream(self.xfer)
        with torch.cuda.device(self.dev):
            self.t._lib._check(self.t.lib().tkr_peer_status(self.buf.peers_ptr, self.flag_off, main.cuda_stream))

    def close(self):
        self.next_state = None
        self.state = None
        self.buf.close()


class DataParallelBpr:
    """Synchronous data-parallel BPR step over the ranks of ``group``.

    ``exchange='peer'`` (default on more than one rank): ``tkr_bpr_dp_step`` -- the item side (V, b and the double-
    buffered gradient accumulators) lives in a peer-mapped exchange buffer; the gradient exchange is fused into the
    update kernel (reduce-scatter by NVLink loads, one RMSProp update per owned item row, all-gather by NVLink stores,
    two flag barriers), overlapped with this rank's user-row updates.  ``state['V']`` / ``state['b']`` are re-pointed
    at views of that buffer; msV / msb rows are current only on their owner (row % world) until ``sync_slots()``.
    ``exchange='nccl'``: tkr_bpr_grad -> ``all_reduce`` of [GV|Gb|tchV] -> tkr_bpr_apply (the round-1 route, kept as
    the baseline the fused kernel is measured against)."""

    def __init__(self, cfg, state, batch, group=None, exchange=None):
        import topkrec
        self.t = topkrec
        self.cfg, self.st, self.batch, self.group = cfg, state, int(batch), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.exchange = exchange or ("peer" if self.world > 1 else "local")
        dev = state["U"].device
        self.ws = topkrec.bpr_workspace(cfg, batch, dev)
        self.item_grads = topkrec.bpr_item_grad_view(cfg, batch, self.ws)
        self.epoch = 0
        self.buf = None
        if self.exchange == "peer":
            from .peer import PeerBuffer
            lay = topkrec.bpr_dp_layout(cfg)
            with torch.cuda.device(dev):
                self.buf = PeerBuffer(lay["total"], group, dev)
            V = self.buf.view(lay["V"], (cfg.c.n_items, cfg.c.d))
            b = self.buf.view(lay["b"], (cfg.c.n_items,))
            V.copy_(state["V"]); b.copy_(state["b"])
            state["V"], state["b"] = V, b              # the replicas now live in the exchange buffer
            torch.cuda.synchronize(dev)
            if self.world > 1:
                dist.barrier(group=group)              # every rank's buffer is initialised before anyone steps

    def step(self, u=None, i=None, j=None, sampler=None, first_draw=0, loss=None):
        st, t = self.st, self.t
        if self.exchange == "peer":
            self.epoch += 1
            t.bpr_dp_step(self.cfg, st["U"], st["msU"], st["msV"], st["msb"], u, i, j, self.batch, self.ws, self.buf, self.epoch,
                          loss, sampler, first_draw)
            return
        dp = self.world > 1
        t.bpr_grad(self.cfg, st["U"], st["V"], st["b"], u, i, j, self.batch, self.ws, loss, sampler, first_draw, data_parallel=dp)
        if dp:
            dist.all_reduce(self.item_grads, op=dist.ReduceOp.SUM, group=self.group)
        t.bpr_apply(self.cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], self.batch, self.ws, data_parallel=dp)

    def check(self):
        """synchronise and raise if a cross-GPU barrier of the fused kernel timed out"""
        if self.exchange == "peer":
            self.t.bpr_dp_status(self.cfg, self.buf)
        else:
            torch.cuda.synchronize(self.st["U"].device)

    def sync_slots(self):
        """fused route: bring every rank's msV / msb up to date (each row from its owner), e.g. before a checkpoint"""
        if self.exchange != "peer" or self.world == 1:
            return
        for name in ("msV", "msb"):
            x = self.st[name]
            mine = torch.zeros_like(x)
            mine[self.rank::self.world] = x[self.rank::self.world]
            dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=self.group)
            x.copy_(mine)

    def close(self):
        if self.buf is not None:
            V, b = self.st["V"].clone(), self.st["b"].clone()
            self.st["V"], self.st["b"] = V, b
            self.buf.close()
            self.buf = None


class DataParallelVbpr:
    """Synchronous data-parallel VBPR step: users partitioned over the ranks (U = [ur|uc] rows never
    move), every other table replicated.  Per step: tkr_vbpr_grad (projection, gradients, this rank's dE / dc) -> all-reduce
    of the item-side region [GV|Gb|tchV] (which carries the content gradient W in its F.E columns) and of [GE|Gc] (dE and
    dc are linear in W: the sum of the local products is the product of the sum, and NCCL hands every rank the same bits)
    -> tkr_vbpr_apply.  Replicas of V, rb, E, c stay bit-identical; equals one GPU stepping the union batch up to fp32
    summation order."""

    def __init__(self, cfg, state, F, batch, group=None):
        import topkrec
        self.t, self.cfg, self.st, self.F, self.batch, self.group = topkrec, cfg, state, F, int(batch), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ws = topkrec.vbpr_workspace(cfg, batch, F.device)
        self.sparse, self.dense = topkrec.vbpr_grad_views(cfg, batch, self.ws)

    def step(self, u=None, i=None, j=None, sampler=None, first_draw=0, loss=None):
        dp = self.world > 1
        self.t.vbpr_grad(self.cfg, self.st, self.F, u, i, j, self.batch, self.ws, loss, sampler, first_draw, data_parallel=dp)
        if dp:
            dist.all_reduce(self.sparse, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.dense, op=dist.ReduceOp.SUM, group=self.group)
        self.t.vbpr_apply(self.cfg, self.st, self.batch, self.ws, None, data_parallel=dp)

    def finish(self):
        """refresh V[:, k/2:] = F.E and bsum = rb + F.c from the final E, c (the export layout, vbpr.py:124-126)"""
        self.t.vbpr_project(self.cfg, self.st, self.F)


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

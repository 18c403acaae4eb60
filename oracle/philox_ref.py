"""numpy restatement of the device triple sampler (``tkr_bpr_sample``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Integer work: the CUDA
sampler must match this bit for bit.

Distribution = ``single/bpr.py:155-165`` of the reference (user uniform over
``tr_users`` with replacement, positive uniform over the user's positives,
negative uniform over all items, redrawn while it is one of the user's
positives).  The *stream* is ours: Philox4x32-10 (Salmon et al., SC'11) with
counter (draw_lo, draw_hi, round, 0) and key (seed_lo, seed_hi); a 32-bit word r
maps to [0, n) as (r * n) >> 32.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
MAX_ROUNDS = 1024      # kMaxSampleRounds of the CUDA sampler


def philox4x32(counter, seed):
    """counter: uint32 [n,4]; returns uint32 [n,4]."""
    c = [counter[:, t].astype(np.uint64) for t in range(4)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [h1 ^ c[1] ^ np.uint64(k0), l1, h0 ^ c[3] ^ np.uint64(k1), l0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def bounded(r, n):
    return ((r.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def sample(tr_users, pos_indptr, pos_idx, n_items, seed, first_draw, n):
    """Triples for draws first_draw .. first_draw+n-1 (pos_idx ascending per user)."""
    tr_users = np.asarray(tr_users); pos_indptr = np.asarray(pos_indptr); pos_idx = np.asarray(pos_idx)
    draws = np.arange(first_draw, first_draw + n, dtype=np.uint64)
    ctr = np.zeros((n, 4), np.uint32)
    ctr[:, 0] = (draws & MASK).astype(np.uint32); ctr[:, 1] = (draws >> np.uint64(32)).astype(np.uint32)
    r = philox4x32(ctr, seed)
    u = tr_users[bounded(r[:, 0], len(tr_users))].astype(np.int64)
    beg = pos_indptr[u]; cnt = pos_indptr[u + 1] - beg
    i = pos_idx[beg + ((r[:, 1].astype(np.uint64) * cnt.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)]
    j = np.empty(n, np.int64)
    for t in range(n):
        pos = set(pos_idx[beg[t]:beg[t] + cnt[t]].tolist())
        cands = [int(bounded(r[t:t + 1, 2], n_items)[0]), int(bounded(r[t:t + 1, 3], n_items)[0])]
        pick = next((c for c in cands if c not in pos), None)
        rnd = 1
        last = cands[-1]
        while pick is None and rnd < MAX_ROUNDS:
            cc = ctr[t:t + 1].copy(); cc[0, 2] = rnd
            rr = philox4x32(cc, seed)[0]
            for w in rr:
                last = int(bounded(np.array([w], np.uint32), n_items)[0])
                if last not in pos:
                    pick = last
                    break
            rnd += 1
        j[t] = pick if pick is not None else last
    return u.astype(np.int32), i.astype(np.int32), j.astype(np.int32)

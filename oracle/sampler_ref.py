"""Replay of the reference's triple sampler and training-file loader.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pinned against the
reference's own ``BPR.load_training_data`` / ``BPR._uniform_user_sampling``
(imported through ``oracle/tf_stub``) by ``tests/golden/make_golden.py``.

Follows ``single/bpr.py:155-165`` (sampler), ``single/bpr.py:51-69,167-171`` and
``utils.py:10-16,58-70`` (loader).  RNG call order, on the
legacy global ``np.random`` stream:

    ub    <- choice(tr_users, B)            == tr_users[randint(0, n_tr, B)]
    for n: ib[n] <- choice(pos[ub[n]])      == pos[randint(0, len(pos))]
           jb[n] <- choice(n_items)         == randint(0, n_items)
           while jb[n] in pos[ub[n]]: redraw
"""
from __future__ import annotations

import numpy as np


def load_ids(path):
    """``utils.py:10-16``: id string -> line index."""
    ids = {}
    with open(path) as f:
        for line in f:
            ids[line.strip()] = len(ids)
    return ids


def load_positives(tr_file, uids, iids):
    """(``tr_users``, ``tr_data``) exactly as ``BPR.load_training_data`` builds
    them: positives only (``like == '1'``), users in first-appearance order,
    items in file order (``utils.py:58-70``, ``bpr.py:167-171``)."""
    tr_data = {}
    n_pos = 0
    with open(tr_file) as f:
        for line in f:
            terms = line.strip().split(",")
            if terms[0] not in uids or len(terms) < 2:
                continue
            urow = uids[terms[0]]
            for t in terms[1:]:
                iid, like = t.split(":")[0], t.split(":")[1]
                if iid in iids and like == "1":
                    tr_data.setdefault(urow, []).append(iids[iid])
                    n_pos += 1
    return list(tr_data.keys()), tr_data, n_pos


def replay_sampler(tr_users, tr_data, n_items, batch_size, n_batches, rs):
    """First ``n_batches`` yields of ``_uniform_user_sampling`` drawn from the
    legacy ``RandomState`` ``rs`` (pass ``np.random`` after ``np.random.seed``
    to replay the global stream).  Returns fresh int32 arrays [n_batches, B]
    (the reference reuses its ``ib``/``jb`` buffers across yields, D-7)."""
    tr_users_arr = np.asarray(tr_users)
    pos_sets = {u: set(v) for u, v in tr_data.items()}
    ub = np.zeros((n_batches, batch_size), np.int32)
    ib = np.zeros((n_batches, batch_size), np.int32)
    jb = np.zeros((n_batches, batch_size), np.int32)
    for t in range(n_batches):
        users = tr_users_arr[rs.randint(0, len(tr_users_arr), batch_size)]
        ub[t] = users
        for n in range(batch_size):
            pos = tr_data[int(users[n])]
            ib[t, n] = pos[rs.randint(0, len(pos))]
            neg = rs.randint(0, n_items)
            while neg in pos_sets[int(users[n])]:
                neg = rs.randint(0, n_items)
            jb[t, n] = neg
    return ub, ib, jb


def to_csr(tr_users, tr_data, n_users):
    """Positives as CSR over *all* user rows (empty for users w/o positives)."""
    indptr = np.zeros(n_users + 1, np.int64)
    for u in tr_users:
        indptr[u + 1] = len(tr_data[u])
    np.cumsum(indptr, out=indptr)
    idx = np.zeros(int(indptr[-1]), np.int32)
    for u in tr_users:
        idx[indptr[u]:indptr[u + 1]] = tr_data[u]
    return indptr, idx

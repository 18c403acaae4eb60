"""Restatement of the reference evaluator (``evaluate.py:57-117``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pinned: the unmodified
``evaluate.py`` was run on ``tests/golden/mini`` (and, in the build container,
on the shipped fold 0) and its printed numbers are committed under
``tests/golden``; this module must reproduce them.

Unlike the script it exposes the per-user filtered top-``total`` lists, which is
what the CUDA path is compared against row by row.
"""
from __future__ import annotations

import os

import numpy as np

from . import topk_ref
from .sampler_ref import load_ids


def load_dat(path, n_rows=None):
    """``evaluate.py:19-28``: text rows -> fp32 matrix (row r = id on line r)."""
    with open(path) as f:
        rows = [ln.strip().split(" ") for ln in f]
    m = np.array([[np.float32(t) for t in r] for r in rows], np.float32)
    assert n_rows is None or m.shape[0] >= n_rows
    return m


def load_history(path):
    """``evaluate.py:30-45``: uid -> set of ALL item ids on the user's line."""
    rated = {}
    with open(path) as f:
        for line in f:
            terms = line.strip().split(",")
            rated[terms[0]] = {t.split(":")[0] for t in terms[1:]}
    return rated


def rated_csr(uids, rated, teids):
    """rated sets -> CSR over user rows of *sorted te-column indices*."""
    n_users = len(uids)
    per_row = [()] * n_users
    for uid, items in rated.items():
        if uid in uids:
            per_row[uids[uid]] = sorted(teids[v] for v in items if v in teids)
    indptr = np.zeros(n_users + 1, np.int64)
    indptr[1:] = np.cumsum([len(r) for r in per_row])
    idx = np.fromiter((c for r in per_row for c in r), np.int32, count=int(indptr[-1]))
    return indptr, idx


def hits_from_lists(lists, uids, teids, te_file, step, total):
    """``evaluate.py:84-112`` given each user's filtered top-``total`` columns."""
    interval = total // step
    tres = np.zeros(interval, np.float64)
    tcount = 0
    with open(te_file) as f:
        for line in f:
            terms = line.strip().split(",")
            likes = {teids[t.split(":")[0]] for t in terms[1:] if int(t.split(":")[1]) == 1}
            if not likes:
                continue
            row = lists[uids[terms[0]]]
            for p in range(total):
                if row[p] >= 0 and int(row[p]) in likes:
                    tres[p // step:] += 1
            tcount += len(likes)
    return tres, tcount


def evaluate(data_dir, model_dir, fold=0, step=5, total=30, scenarios=("im",), scorer="fma", use_bias=True):
    """Returns {scenario: (acc[interval], lists int32 [n_users,total])}.

    scorer='fma'  : the parity definition (``oracle/topk_ref.c``)
    scorer='blas' : np.dot + stable argsort (the reference's literal calls)
    Bias follows the intended per-column gather (``old/methods/bpr_test.py:18-32``);
    the shipped line ``evaluate.py:80`` only works when n_te == n_items.
    """
    uids = load_ids(os.path.join(data_dir, "uid"))
    vids = load_ids(os.path.join(data_dir, "vid"))
    rated = load_history(os.path.join(data_dir, "f%dtr.txt" % fold))
    U = load_dat(os.path.join(model_dir, "final-U.dat"), len(uids))
    V = load_dat(os.path.join(model_dir, "final-V.dat"), len(vids))
    bpath = os.path.join(model_dir, "final-B.dat")
    Bv = load_dat(bpath, len(vids)).ravel() if use_bias and os.path.exists(bpath) else None
    out = {}
    for sc in scenarios:
        teids = load_ids(os.path.join(data_dir, "f%dte.%s.idl" % (fold, sc)))
        cols = np.array([vids[v] for v in teids], np.int64)
        Vte = V[cols]
        bias = Bv[cols] if Bv is not None else None
        indptr, idx = rated_csr(uids, rated, teids)
        fn = topk_ref.score_topk if scorer == "fma" else topk_ref.score_topk_numpy
        lists, _ = fn(U, Vte, total, bias, indptr, idx)
        tres, tcount = hits_from_lists(lists, uids, teids, os.path.join(data_dir, "f%dte.%s.txt" % (fold, sc)), step, total)
        out[sc] = (tres / tcount, lists)
    return out


def format_line(sc, acc):
    """``evaluate.py:113-117``."""
    return sc + "".join(",%.6f" % a for a in acc)

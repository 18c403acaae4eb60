"""Python face of the C oracle for hot path 2 (``oracle/topk_ref.c``) plus a
numpy restatement of the reference's literal ``np.dot`` + ``np.argsort`` route.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows ``evaluate.py:75-105``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import clib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def score_topk(U, V, k, bias=None, rated_indptr=None, rated_idx=None, col_offset=0):
    """Exact filtered top-k per user: fp32 FMA-chain scores, order (score desc,
    column desc), rated columns skipped.  Returns (idx int32 [nu,k], score fp32)."""
    U = np.ascontiguousarray(U, np.float32); V = np.ascontiguousarray(V, np.float32)
    nu, d = U.shape; ni = V.shape[0]
    assert V.shape[1] == d
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32).ravel(); assert bias.shape[0] == ni
    if rated_indptr is not None:
        rated_indptr = np.ascontiguousarray(rated_indptr, np.int64)
        rated_idx = np.ascontiguousarray(rated_idx, np.int32)
    out_idx = np.empty((nu, k), np.int32); out_score = np.empty((nu, k), np.float32)
    rc = clib.lib().tkr_ref_score_topk(
        _p(U, ctypes.c_float), ctypes.c_int64(nu), _p(V, ctypes.c_float), ctypes.c_int64(ni), ctypes.c_int(d),
        _p(bias, ctypes.c_float), _p(rated_indptr, ctypes.c_int64), _p(rated_idx, ctypes.c_int32),
        ctypes.c_int(k), ctypes.c_int64(col_offset), _p(out_idx, ctypes.c_int32), _p(out_score, ctypes.c_float))
    assert rc == 0
    return out_idx, out_score


def topk_merge(idx, score):
    """Merge per-shard lists [G, nu, k] -> [nu, k] with the same ordering."""
    idx = np.ascontiguousarray(idx, np.int32); score = np.ascontiguousarray(score, np.float32)
    G, nu, k = idx.shape
    out_idx = np.empty((nu, k), np.int32); out_score = np.empty((nu, k), np.float32)
    rc = clib.lib().tkr_ref_topk_merge(_p(idx, ctypes.c_int32), _p(score, ctypes.c_float), ctypes.c_int(G),
                                       ctypes.c_int64(nu), ctypes.c_int(k), _p(out_idx, ctypes.c_int32),
                                       _p(out_score, ctypes.c_float))
    assert rc == 0
    return out_idx, out_score


def score_topk_numpy(U, V, k, bias=None, rated_indptr=None, rated_idx=None, stable=True):
    """The reference's literal route: ``np.dot`` (BLAS order) then full argsort
    read backwards (``evaluate.py:78,81,96-105``).  ``stable=False`` is the
    reference's default (unstable) sort.  Used to quantify how many rows differ
    from the FMA-chain definition only because of ties / near-ties."""
    S = np.dot(np.asarray(U, np.float32), np.asarray(V, np.float32).T)
    if bias is not None:
        S += np.asarray(bias, np.float32).reshape(1, -1)
    order = np.argsort(S, axis=1, kind="stable" if stable else None)[:, ::-1]
    nu = U.shape[0]
    out_idx = np.full((nu, k), -1, np.int32); out_score = np.full((nu, k), -np.inf, np.float32)
    for r in range(nu):
        cols = order[r]
        if rated_indptr is not None:
            rated = rated_idx[rated_indptr[r]:rated_indptr[r + 1]]
            if rated.size:
                cols = cols[~np.isin(cols, rated)]
        cols = cols[:k]
        out_idx[r, :cols.size] = cols
        out_score[r, :cols.size] = S[r, cols]
    return out_idx, out_score

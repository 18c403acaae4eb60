"""Host side of the ALS half-step (``tkr_als_solve_rows`` / ``tkr_als_gram``, include/topkrec.h): the work list
that cuts each row's positives into segments, and thin wrappers.  No CPU path: tensors must be on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import TkrError, lib, _check, _dev, _stream, _need_cuda, tkr_als_cfg, tkr_als_plan


def build_plan(indptr, seg):
    """Segment work list of ``tkr_als_plan`` (host arrays): rows longest first, each cut into ``ceil(n / seg)``
    segments (empty rows: one empty segment); rows with more than one segment get consecutive partial slots."""
    indptr = np.ascontiguousarray(indptr, np.int64)
    if indptr.ndim != 1 or indptr.size < 2 or indptr[0] != 0 or np.any(np.diff(indptr) < 0):
        raise ValueError("indptr is not a CSR row pointer (need >= 1 row)")
    if seg < 1:
        raise ValueError("seg must be >= 1")
    cnt = np.diff(indptr)
    nseg = np.maximum(1, -(-cnt // seg))
    order = np.argsort(-cnt, kind="stable")
    rows = np.repeat(order, nseg[order])
    first = np.r_[0, np.cumsum(nseg[order])[:-1]]
    within = np.arange(rows.size) - np.repeat(first, nseg[order])
    seg_off = indptr[rows] + within * seg
    seg_len = np.minimum(seg, indptr[rows + 1] - seg_off).astype(np.int32)
    multi = order[nseg[order] > 1]
    is_multi = nseg[rows] > 1
    seg_slot = np.full(rows.size, -1, np.int32)
    seg_slot[is_multi] = np.arange(int(is_multi.sum()), dtype=np.int32)
    slot0 = np.r_[0, np.cumsum(nseg[multi])[:-1]] if multi.size else np.zeros(0, np.int64)
    return dict(seg_row=rows.astype(np.int32), seg_off=seg_off.astype(np.int64), seg_len=seg_len, seg_slot=seg_slot,
                multi_row=multi.astype(np.int32), multi_slot0=np.asarray(slot0, np.int32),
                multi_nslots=nseg[multi].astype(np.int32), multi_total=cnt[multi].astype(np.int64)), int(is_multi.sum())


class AlsSide:
    """One side of the alternation (users solved from items, or items from users): the CSR of positives of the
    solved rows on the device plus the segment work list.

    indptr int64[n_rows+1], idx int32[nnz] (host): positives of each solved row, as rows of the fixed side
    (``usm`` / ``ism`` of wmf.py:35-52, duplicates kept).  ``seg`` = most positives one thread block accumulates;
    longer rows are split into partial sums (deterministic: summed in segment order).  The default also bounds the length of
    one sequential fp32 accumulation: on the reference's uniform(0,1) start a 5000-positive row solved from one 4096-long
    chain sits 6.8e-5 from the fp64-exact solution, from 1024-long chains 1.5e-5 (BLAS order: 0.8e-5;
    profiles/tf32_gram_model.py)."""

    def __init__(self, indptr, idx, seg=1024, device="cuda"):
        indptr = np.ascontiguousarray(indptr, np.int64)
        idx = np.ascontiguousarray(idx, np.int32)
        self.host, self.n_slots = build_plan(indptr, seg)
        if indptr[-1] != idx.size:
            raise ValueError("indptr[-1] != len(idx)")
        self.n_rows = indptr.size - 1
        self.nnz = int(idx.size)
        if idx.size and int(idx.min()) < 0:
            raise ValueError("idx holds a negative row index (%d)" % int(idx.min()))
        self.max_idx = int(idx.max()) if idx.size else -1
        self.seg = int(seg)
        self.rated = np.flatnonzero(np.diff(indptr) > 0).astype(np.int32)       # u_rated / i_rated (wmf.py:53-54), ascending
        _need_cuda()
        dev = torch.device(device)
        self.dev = {k: torch.from_numpy(v).to(dev) for k, v in self.host.items()}
        self.idx = torch.from_numpy(idx).to(dev)
        self.rated_dev = torch.from_numpy(self.rated).to(dev)
        self.plan = tkr_als_plan(self.host["seg_row"].size, *(self.dev[k].data_ptr() for k in ("seg_row", "seg_off", "seg_len", "seg_slot")),
                                 self.host["multi_row"].size, *(self.dev[k].data_ptr() for k in ("multi_row", "multi_slot0", "multi_nslots", "multi_total")),
                                 self.n_slots)
        self._partial = None
        self.device = dev

    def partial(self, d):
        need = lib().tkr_als_partial_bytes(int(d), self.n_slots)
        if self.n_slots == 0:
            return None, 0
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._partial, need


_gram_ws = {}


def als_gram(Y, rows, scale, ridge, out=None):
    """``scale * Y[rows].T @ Y[rows] + ridge * I`` (cer.py:37-38, :47-48) as fp32 [d,d] on the device."""
    _need_cuda(Y, rows)
    d = Y.shape[1]
    if out is None:
        out = torch.empty(d, d, dtype=torch.float32, device=Y.device)
    need = lib().tkr_als_gram_workspace_bytes(d)
    if need == 0:
        raise TkrError("als_gram: d=%d is outside [1, 256]" % d)
    key = (Y.device, d)
    if key not in _gram_ws:
        _gram_ws[key] = torch.empty(need, dtype=torch.uint8, device=Y.device)
    ws = _gram_ws[key]
    _check(lib().tkr_als_gram(_dev(Y, torch.float32, "Y"), d, _dev(rows, torch.int32, "rows"), rows.numel(), scale, ridge,
                              _dev(out, torch.float32, "out"), ws.data_ptr(), ws.numel(), _stream()))
    return out


def als_solve_rows(side: AlsSide, Y, X, base, a, b, ridge, lreg, prior=None, solve_empty=False, item_loss=False, loss_rows=None):
    """Solve every row of ``X`` from the fixed factor ``Y`` (one half-step; cer.py:39-45 / :49-62).  Returns the per-row
    loss terms (float64[n_rows], device)."""
    _need_cuda(Y, X, base)
    d = X.shape[1]
    if Y.shape[1] != d or tuple(base.shape) != (d, d) or X.shape[0] != side.n_rows:
        raise ValueError("shape mismatch: X %s, Y %s, base %s, rows %d" % (tuple(X.shape), tuple(Y.shape), tuple(base.shape), side.n_rows))
    if side.max_idx >= Y.shape[0]:
        raise ValueError("idx refers to a row beyond Y")
    if loss_rows is None:
        loss_rows = torch.empty(side.n_rows, dtype=torch.float64, device=X.device)
    cfg = tkr_als_cfg(d, a, b, ridge, lreg, int(bool(solve_empty)), int(bool(item_loss)))
    part, nbytes = side.partial(d)
    _check(lib().tkr_als_solve_rows(C.byref(cfg), C.byref(side.plan), _dev(Y, torch.float32, "Y"), _dev(X, torch.float32, "X"),
                                    side.idx.data_ptr(), _dev(base, torch.float32, "base"), _dev(prior, torch.float32, "prior"),
                                    _dev(loss_rows, torch.float64, "loss_rows"), part.data_ptr() if part is not None else None,
                                    nbytes, _stream()))
    return loss_rows

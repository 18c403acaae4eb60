#!/usr/bin/env python3
"""Accelerated evaluator with the reference's CLI and output (``evaluate.py`` of
domainxz/top-k-rec): same flags, same ``<scenario>,a@5,...`` lines.

The reference's ``np.dot`` (:78) + ``np.argsort`` (:81) + Python rated-filter walk
(:96-105) become one fused device call per user batch (``tkr_score_topk``): scores
never leave the SM, rated items are masked from a CSR, and only the filtered
top-``total`` columns per user stay on the device, where ``tkr_eval_hits`` counts the hits
(``evaluate.py:84-112``); only ``total`` counters come back.
Bias is gathered per test column (the intent of ``old/methods/bpr_test.py:18-32``;
the shipped ``evaluate.py:80`` broadcast only works when n_te == n_items).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from utils import get_id_dict_from_file, get_embed_from_file, rated_csr_from_files, test_lines_from_files  # noqa: E402


MAX_TOTAL = 64     # tkr_score_topk: k <= 64


def filtered_topk(umat, temat, total, bias, rated_indptr, rated_idx, user_batch=65536):
    """Per-user filtered top-``total`` test columns via the device engine (device tensor [n_users, total]); user
    batches are uploaded on a side stream while the previous batch is scored (``topkrec.score_topk_batches``)."""
    import torch
    import topkrec
    dev = torch.device('cuda')
    V = torch.from_numpy(temat).to(dev)
    b = torch.from_numpy(bias).to(dev) if bias is not None else None
    ridx = torch.from_numpy(rated_idx).to(dev)
    lists, _ = topkrec.score_topk_batches(umat, V, total, b, rated_indptr, ridx, user_batch=user_batch, engine='tc', to_host=False)
    return lists


def count_hits(lists, uid_file, te_file, te_idl_file, step, total):
    """``evaluate.py:84-112``: hits@{step, 2*step, ...} over test lines with >= 1 like, counted on the device."""
    import torch
    import topkrec
    rows, indptr, idx = test_lines_from_files(uid_file, te_file, te_idl_file)
    dev = lists.device
    hits, _ = topkrec.eval_hits(lists, torch.from_numpy(rows).to(dev), torch.from_numpy(indptr).to(dev),
                                torch.from_numpy(idx).to(dev), step)
    return hits, int(indptr[-1])


def main(argv=None):
    parser = argparse.ArgumentParser(description='Evaluate weighted matrix factorization based methods.')
    parser.add_argument('-d', '--data', required=True, help='The data path for the evaluation')
    parser.add_argument('-m', '--model', required=True, help='The work path for the model')
    parser.add_argument('-f', '--fold', type=int, default=0, help='The index of evaluation fold')
    parser.add_argument('-s', '--step', type=int, default=5, help='The number of evaluation step')
    parser.add_argument('-t', '--total', type=int, default=30, help='The number of total predictions')
    parser.add_argument('-sl', '--scenarios', nargs='+', default=None, help='The test scenario list')
    args = parser.parse_args(argv)
    # the device engine keeps the filtered top-`total` of a user in registers / shared memory: total <= 64
    # (the reference accepts any total up to the number of test items; its own recipe uses 30)
    if not 1 <= args.total <= MAX_TOTAL:
        parser.error('--total must be in [1, %d] on the device engine (got %d)' % (MAX_TOTAL, args.total))
    if args.step < 1:
        parser.error('--step must be positive')

    lines, _ = run(args.data, args.model, args.fold, args.step, args.total, args.scenarios)
    for line in lines:
        print(line)
    return lines


def run(data, model, fold=0, step=5, total=30, scenarios=('im',), use_bias=True, keep_lists=False):
    """The body of the reference's ``__main__`` (``evaluate.py:57-117``) on the device engine.  Returns the printed
    lines and, with ``keep_lists``, {scenario: int32 [n_users, total] filtered top-``total`` test columns} (host)."""
    uid_file, tr_file = os.path.join(data, 'uid'), os.path.join(data, 'f%dtr.txt' % fold)
    uids = get_id_dict_from_file(uid_file)
    vids = get_id_dict_from_file(os.path.join(data, 'vid'))
    umat = get_embed_from_file(os.path.join(model, 'final-U.dat'), uids)
    vmat = get_embed_from_file(os.path.join(model, 'final-V.dat'), vids)
    bmat = get_embed_from_file(os.path.join(model, 'final-B.dat'), vids) if use_bias else None
    lines, kept = [], {}
    for sc in scenarios:
        te_idl = os.path.join(data, 'f%dte.%s.idl' % (fold, sc))
        teids = get_id_dict_from_file(te_idl)
        cols = np.fromiter((vids[v] for v in teids), np.int64, count=len(teids))
        temat = np.ascontiguousarray(vmat[cols])
        bias = np.ascontiguousarray(bmat.ravel()[cols]) if bmat is not None else None
        indptr, idx = rated_csr_from_files(uid_file, tr_file, te_idl, len(uids))
        lists = filtered_topk(umat, temat, total, bias, indptr, idx)
        hits, tcount = count_hits(lists, uid_file, os.path.join(data, 'f%dte.%s.txt' % (fold, sc)), te_idl, step, total)
        lines.append(sc + ''.join(',%.6f' % (h / tcount) for h in hits))
        if keep_lists:
            kept[sc] = lists.cpu().numpy()
    return lines, kept


if __name__ == '__main__':
    main()

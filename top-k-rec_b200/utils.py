"""Host utilities with the reference's names and file formats (``utils.py`` of
domainxz/top-k-rec), rewritten around whole-file numpy operations.

Formats (README.md:56-69 of the reference):
  id file      one opaque id string per line; row index = line number
  rating file  ``uid,iid:like,iid:like,...`` with like in {0,1}
  .dat         text matrix, every element ``'%f '`` (6 decimals + one space),
               one row per line (reference writer ``utils.py:47-55``)
"""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np


def tprint(msg: str) -> None:
    """Timestamped print, same layout as the reference (``utils.py:6-7``)."""
    print('%s: %s' % (datetime.now().strftime('%Y-%m-%d %H:%M:%S.%f'), msg))


def _is_file(path) -> bool:
    return os.path.isfile(path)


def get_id_dict_from_file(file_path: str) -> dict:
    """id string -> 0-based line index; a missing file gives ``{}`` (``utils.py:10-16``)."""
    if not _is_file(file_path):
        return dict()
    with open(file_path, 'r') as f:
        return {line.strip(): n for n, line in enumerate(f)}


def get_iv_dict_from_file(file_path: str) -> dict:
    """line index -> id string (``utils.py:19-25``)."""
    if not _is_file(file_path):
        return dict()
    with open(file_path) as f:
        return dict(enumerate(line.strip() for line in f))


def get_embed_from_file(file_path: str, ids: dict = None):
    """Read a ``.dat`` matrix as fp32 (``utils.py:28-44``).  With ``ids`` the
    result has ``len(ids)`` rows and row ``ids[x]`` comes from line ``ids[x]``."""
    if not _is_file(file_path):
        return None
    with open(file_path) as f:
        text = f.read()
    first = text.find('\n')
    n_col = len(text[:first if first >= 0 else len(text)].split())
    # tokens -> double -> fp32: the same two roundings as np.float32(str)
    flat = np.array(text.split(), dtype=np.float64)
    mat = flat.reshape(-1, n_col).astype(np.float32)
    if ids is not None:
        rows = np.fromiter(ids.values(), np.int64, count=len(ids))
        out = np.zeros((len(ids), n_col), np.float32)
        out[rows] = mat[rows]
        return out
    return mat


def export_embed_to_file(file_path: str, embed) -> None:
    """Write a ``.dat`` matrix, byte-identical to the reference writer
    (``utils.py:47-55``): ``'%f '`` per element, newline per row."""
    parent = os.path.dirname(file_path)
    if parent and not os.path.isdir(parent):
        os.mkdir(parent)
    embed = np.asarray(embed)
    if embed.ndim != 2:
        raise ValueError('embed must be a matrix, got shape %s' % (embed.shape,))
    fmt = '%f ' * embed.shape[1] + '\n'
    rows = embed.astype(np.float64)          # exact: what '%f' % np.float32 formats
    with open(file_path, 'w') as f:
        f.writelines(fmt % tuple(r) for r in rows)


def _iter_ratings(file_path):
    """Yield (uid, [(iid, like_str), ...]) per line of a rating file."""
    with open(file_path, 'r') as f:
        for line in f:
            terms = line.strip().split(',')
            yield terms[0], [tuple(t.split(':')[:2]) for t in terms[1:]]


def get_data_from_file(file_path: str, uids: dict, iids: dict) -> list:
    """Positive (uid, iid) pairs in file order (``utils.py:58-70``): known user,
    known item, like == '1'."""
    data = list()
    if not _is_file(file_path):
        return data
    for uid, pairs in _iter_ratings(file_path):
        if uid in uids and pairs:
            data.extend((uid, iid) for iid, like in pairs if like == '1' and iid in iids)
    return data


def get_history_from_file(file_path: str):
    """(browsed, counter): every item on a user's line, and per-item like counts
    (``utils.py:73-89``)."""
    browsed, counter = dict(), dict()
    if not _is_file(file_path):
        return browsed, counter
    for uid, pairs in _iter_ratings(file_path):
        browsed[uid] = {iid for iid, _ in pairs}
        for iid, like in pairs:
            if like == '1':
                counter[iid] = counter.get(iid, 0) + 1
    return browsed, counter


def positives_csr(tr_users, tr_data, n_users):
    """Adjacency lists -> CSR over all user rows, items ascending within a user
    (the layout the device sampler binary-searches)."""
    indptr = np.zeros(n_users + 1, np.int64)
    for u in tr_users:
        indptr[u + 1] = len(tr_data[u])
    np.cumsum(indptr, out=indptr)
    idx = np.empty(int(indptr[-1]), np.int32)
    for u in tr_users:
        idx[indptr[u]:indptr[u + 1]] = np.sort(np.asarray(tr_data[u], np.int32))
    return indptr, idx


def rated_csr(uids: dict, browsed: dict, teids: dict):
    """Per user row, the ascending test-column indices of every item the user has
    rated in training (what ``evaluate.py:98`` filters out)."""
    n_users = len(uids)
    lists = [None] * n_users
    for uid, items in browsed.items():
        r = uids.get(uid)
        if r is not None:
            lists[r] = sorted(teids[v] for v in items if v in teids)
    indptr = np.zeros(n_users + 1, np.int64)
    indptr[1:] = np.cumsum([len(x) if x else 0 for x in lists])
    idx = np.fromiter((c for x in lists if x for c in x), np.int32, count=int(indptr[-1]))
    return indptr, idx

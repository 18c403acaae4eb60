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
    result has ``len(ids)`` rows and row ``ids[x]`` comes from line ``ids[x]``.
    Parsing is the native codec ``tkr_dat_read`` (tokens -> double -> fp32, the
    same two roundings as ``np.float32(str)``)."""
    if not _is_file(file_path):
        return None
    import topkrec
    mat = topkrec.dat_read(file_path)
    if ids is not None:
        rows = np.fromiter(ids.values(), np.int64, count=len(ids))
        out = np.zeros((len(ids), mat.shape[1]), np.float32)
        out[rows] = mat[rows]
        return out
    return mat


def export_embed_to_file(file_path: str, embed) -> None:
    """Write a ``.dat`` matrix, byte-identical to the reference writer
    (``utils.py:47-55``): ``'%f '`` per element, newline per row (``tkr_dat_write``;
    a float64 matrix -- CER's ``E`` -- is formatted from its doubles, ``tkr_dat_write_f64``)."""
    parent = os.path.dirname(file_path)
    if parent and not os.path.isdir(parent):
        os.mkdir(parent)
    embed = np.asarray(embed)
    if embed.ndim != 2:
        raise ValueError('embed must be a matrix, got shape %s' % (embed.shape,))
    import topkrec
    topkrec.dat_write(file_path, embed)


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


# ----------------------------------------------------------------------------- native rating-file paths
def _sorted_unique(key):
    """np.unique for int64 keys via sort + neighbour mask (numpy 2.3's hash-based unique is ~20x slower here)."""
    key = np.sort(key)
    if key.size == 0:
        return key
    return key[np.r_[True, key[1:] != key[:-1]]]


def positives_from_files(uid_file: str, iid_file: str, tr_file: str):
    """The loader of ``bpr.py:51-69`` + ``:167-171`` on the native parser (``tkr_ratings_parse``): returns
    ``(n_pairs, tr_users, tr_data)`` with ``tr_users`` in first-appearance order and ``tr_data[u]`` the user's
    positives in file order -- the same structures ``get_data_from_file`` + ``_data_to_training_dict`` build."""
    import topkrec
    line_user, indptr, item, like = topkrec.ratings_parse(tr_file, uid_file, iid_file)
    users = np.repeat(line_user, np.diff(indptr))
    keep = (users >= 0) & (item >= 0) & (like == 1)
    u, it = users[keep], item[keep]
    order = np.argsort(u, kind='stable')                       # groups by user, file order kept inside a group
    us, its = u[order], it[order]
    starts = np.flatnonzero(np.r_[True, us[1:] != us[:-1]]) if us.size else np.zeros(0, np.int64)
    first_pos = order[starts] if us.size else starts           # position of each user's first positive in the file
    by_first = np.argsort(first_pos, kind='stable')
    ends = np.r_[starts[1:], us.size]
    tr_data = {int(us[starts[g]]): its[starts[g]:ends[g]].tolist() for g in by_first}
    return int(u.size), list(tr_data.keys()), tr_data


def rated_csr_from_files(uid_file: str, tr_file: str, te_idl_file: str, n_users: int):
    """``rated_csr`` straight from the files (``evaluate.py:30-45,66,98``): per user row the ascending test columns
    of every item on the user's LAST training line (the reference's dict keeps the last occurrence of a uid)."""
    import topkrec
    line_user, indptr, col, _ = topkrec.ratings_parse(tr_file, uid_file, te_idl_file)
    n_lines = line_user.size
    last = np.full(n_users, -1, np.int64)
    known = line_user >= 0
    last[line_user[known]] = np.flatnonzero(known)                # repeated uid: the later (larger) line index is kept
    line_of_pair = np.repeat(np.arange(n_lines), np.diff(indptr))
    users = np.repeat(line_user, np.diff(indptr))
    keep = (users >= 0) & (col >= 0)
    keep[keep] &= last[users[keep]] == line_of_pair[keep]
    key = _sorted_unique(users[keep].astype(np.int64) * (int(col.max()) + 2 if col.size else 1) + col[keep])
    mod = int(col.max()) + 2 if col.size else 1
    rows, cols = key // mod, (key % mod).astype(np.int32)
    out_ptr = np.zeros(n_users + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=out_ptr[1:])
    return out_ptr, cols


def test_lines_from_files(uid_file: str, te_file: str, te_idl_file: str):
    """The test file as the CSR ``tkr_eval_hits`` takes (``evaluate.py:84-93``): per line with >= 1 like its user row
    and its distinct liked test columns, ascending."""
    import topkrec
    line_user, indptr, col, like = topkrec.ratings_parse(te_file, uid_file, te_idl_file)
    line_of_pair = np.repeat(np.arange(line_user.size), np.diff(indptr))
    keep = like == 1
    if np.any(col[keep] < 0) or np.any(line_user[np.unique(line_of_pair[keep])] < 0):
        raise KeyError('test file names an id that is not in the id lists')          # the reference raises KeyError too
    mod = int(col.max()) + 2 if col.size else 1
    key = _sorted_unique(line_of_pair[keep].astype(np.int64) * mod + col[keep])
    lines, cols = key // mod, (key % mod).astype(np.int32)
    ulines, counts = np.unique(lines, return_counts=True)
    out_ptr = np.zeros(ulines.size + 1, np.int64)
    np.cumsum(counts, out=out_ptr[1:])
    return line_user[ulines].astype(np.int32), out_ptr, cols

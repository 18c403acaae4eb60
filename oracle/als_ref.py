"""CPU oracle for the alternating-least-squares path (SURVEY.md 8(f) NEXT-1): TEST INFRASTRUCTURE ONLY.

Restates ``CER.train`` (/root/reference/single/cer.py:24-73) and the *intended* ``WMF.train``
(/root/reference/single/wmf.py:61-101; the shipped method indexes ``self.usm[uid].keys()`` on lists and
cannot run) on flat CSR inputs, with the reference's dtypes: fp32 Gram matrices (``np.dot`` of
fp32 rows), the k x k systems solved by ``np.linalg.solve`` (fp64 LAPACK, result cast to fp32 on assignment),
the loss accumulated in Python floats.

PINNED: ``tests/golden/make_golden_als.py`` runs the UNMODIFIED reference ``CER.train`` (imported through
``oracle/tf_stub``) on tests/golden/mini with a seeded start and commits its factors and printed losses;
``tests/test_oracle.py::test_als_oracle_matches_reference_cer`` requires this file to reproduce them.
"""
import numpy as np


def csr_from_lists(lists, n):
    """dict/list of python lists (``usm`` / ``ism``, wmf.py:35-52) -> (indptr int64[n+1], idx int32[nnz])."""
    indptr = np.zeros(n + 1, np.int64)
    for r in range(n):
        indptr[r + 1] = indptr[r] + len(lists[r])
    idx = np.zeros(int(indptr[-1]), np.int32)
    for r in range(n):
        idx[indptr[r]:indptr[r + 1]] = lists[r]
    return indptr, idx


def shared_gram(Y, rated, b, ridge):
    """``XX = np.dot(Yr.T, Yr) * b + Ik * ridge`` (cer.py:37-38, :47-48): fp32."""
    Yr = Y[np.asarray(rated, np.int64), :]
    k = Y.shape[1]
    return (np.dot(Yr.T, Yr) * b + np.eye(k, dtype=np.float32) * ridge).astype(np.float32)


def user_step(fue, fie, u_indptr, u_idx, i_rated, a, b, lu):
    """cer.py:36-46 == wmf.py:67-77 (unit ratings).  Updates ``fue`` in place, returns the loss term."""
    XX = shared_gram(fie, i_rated, b, lu)
    loss = 0.0
    for i in range(fue.shape[0]):
        pos = u_idx[u_indptr[i]:u_indptr[i + 1]]
        if len(pos) > 0:
            Vi = fie[pos, :]
            fue[i, :] = np.linalg.solve(np.dot(Vi.T, Vi) * (a - b) + XX, np.sum(Vi, axis=0) * a)
        loss += 0.5 * lu * np.sum(fue[i, :] ** 2)
    return float(loss)


def item_step(fue, fie, i_indptr, i_idx, u_rated, a, b, lv, Fe=None):
    """cer.py:47-63 (``Fe`` given: content prior ``F_j E``, unrated items are solved too) or
    wmf.py:78-96 (``Fe is None``: ridge only, unrated items keep their row).  Updates ``fie`` in place."""
    k = fie.shape[1]
    Ik = np.eye(k, dtype=np.float32)
    Ur = fue[np.asarray(u_rated, np.int64), :]
    XX = np.dot(Ur.T, Ur) * b
    loss = 0.0
    for j in range(fie.shape[0]):
        pos = i_idx[i_indptr[j]:i_indptr[j + 1]]
        B = XX.copy()
        if len(pos) > 0:
            Uj = fue[pos, :]
            B += np.dot(Uj.T, Uj) * (a - b)
            rhs = np.sum(Uj, axis=0) * a
            if Fe is not None:
                rhs = rhs + Fe[j, :] * lv
            fie[j, :] = np.linalg.solve(B + Ik * lv, rhs)
            loss += 0.5 * np.linalg.multi_dot((fie[j, :], B, fie[j, :]))
            loss += 0.5 * len(pos) * a
            loss -= np.sum(np.multiply(Uj, fie[j, :])) * a
        elif Fe is not None:
            fie[j, :] = np.linalg.solve(B + Ik * lv, Fe[j, :] * lv)
        if Fe is not None:
            loss += 0.5 * lv * np.sum((fie[j, :] - Fe[j, :]) ** 2)
        else:
            loss += 0.5 * lv * np.sum(fie[j, :] ** 2)
    return float(loss)


def cer_train(fue, fie, E, feat, u_indptr, u_idx, i_indptr, i_idx, a=1.0, b=0.01, lu=0.01, lv=10.0, le=10e3,
              max_iter=200, tol=1e-4):
    """``CER.train`` (cer.py:24-73).  Returns (fue, fie, E, losses); inputs are not modified."""
    fue, fie = fue.copy(), fie.copy()
    n_users, n_items = fue.shape[0], fie.shape[0]
    u_rated = [u for u in range(n_users) if u_indptr[u + 1] > u_indptr[u]]
    i_rated = [j for j in range(n_items) if i_indptr[j + 1] > i_indptr[j]]
    FF = lv * np.dot(feat.T, feat) + le * np.eye(feat.shape[1])
    loss = np.exp(50)
    losses = []
    for _ in range(max_iter):
        Fe = np.dot(feat, E)
        loss_old = loss
        loss = user_step(fue, fie, u_indptr, u_idx, i_rated, a, b, lu)
        loss += item_step(fue, fie, i_indptr, i_idx, u_rated, a, b, lv, Fe)
        E = np.linalg.solve(FF, lv * np.dot(feat.T, fie))
        loss += 0.5 * le * np.sum(E ** 2)
        losses.append(float(loss))
        if np.abs(loss_old - loss) / loss_old < tol:
            break
    Fe = np.dot(feat, E)
    rated = set(i_rated)
    for j in range(n_items):
        if j not in rated:
            fie[j, :] = Fe[j, :]
    return fue, fie, E, losses


def wmf_train(fue, fie, u_indptr, u_idx, i_indptr, i_idx, a=1.0, b=0.01, lu=0.01, lv=0.01, max_iter=200, tol=1e-4):
    """The intended ``WMF.train`` (wmf.py:61-101) with unit ratings."""
    fue, fie = fue.copy(), fie.copy()
    n_users, n_items = fue.shape[0], fie.shape[0]
    u_rated = [u for u in range(n_users) if u_indptr[u + 1] > u_indptr[u]]
    i_rated = [j for j in range(n_items) if i_indptr[j + 1] > i_indptr[j]]
    loss = np.exp(50)
    losses = []
    for _ in range(max_iter):
        loss_old = loss
        loss = user_step(fue, fie, u_indptr, u_idx, i_rated, a, b, lu)
        loss += item_step(fue, fie, i_indptr, i_idx, u_rated, a, b, lv, None)
        losses.append(float(loss))
        if np.abs(loss_old - loss) / loss_old < tol:
            break
    return fue, fie, losses

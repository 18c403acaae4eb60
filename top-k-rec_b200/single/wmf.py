"""Weighted matrix factorisation by alternating least squares on the B200 engine, behind the reference's ``WMF``
surface (``single/wmf.py`` of domainxz/top-k-rec).

The reference's ``train`` (``wmf.py:61-101``) is a Python loop of ``np.dot`` + ``np.linalg.solve`` per user and per
item -- and, as shipped, cannot run (it calls ``.keys()`` on the lists ``load_training_data`` builds; the working
statement of the same alternation is ``CER.train``, ``cer.py:36-63``).  Here one half-step is one call of
``tkr_als_solve_rows``: every row's normal matrix is accumulated, factored and solved by one thread block.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

import topkrec
from utils import tprint, get_id_dict_from_file

from .rec import REC


def _lists(indptr, idx, n):
    parts = np.split(idx, indptr[1:-1]) if n else []
    return {r: parts[r].tolist() for r in range(n)}


class WMF(REC):
    def __init__(self, k: int, lu: float = 0.01, lv: float = 0.01, a: float = 1, b: float = 0.01, device: str = 'cuda',
                 seg: int = 1024) -> None:
        self.__sn = 'wmf'
        self.k = k
        self.lu, self.lv, self.a, self.b = lu, lv, a, b
        self.device, self.seg = device, seg
        self.uids = self.iids = None
        self.n_users = self.n_items = self.n_ratings = None
        self.usm = self.ism = None
        self.u_rated = self.i_rated = None
        self.fue = self.fie = None
        self.losses = []
        self._csr = None
        self._sides = None

    # ------------------------------------------------------------------ data
    def load_training_data(self, uid_file: str, iid_file: str, tr_file: str) -> None:
        """``wmf.py:33-56``: id maps, ``usm`` (user row -> liked item rows, file order), ``ism`` (item row -> users),
        ``u_rated`` / ``i_rated``, and a uniform(0,1) start for both factors."""
        self.uids = get_id_dict_from_file(uid_file)
        self.iids = get_id_dict_from_file(iid_file)
        self.n_users, self.n_items = len(self.uids), len(self.iids)
        line_user, indptr, item, like = topkrec.ratings_parse(tr_file, uid_file, iid_file)
        users = np.repeat(line_user, np.diff(indptr))
        pos = like == 1
        if np.any(users[pos] < 0) or np.any(item[pos] < 0):
            raise KeyError('training file names an id that is not in the id lists')     # wmf.py:51-52 raises KeyError too
        self.set_training_pairs(users[pos], item[pos], self.n_users, self.n_items, lists=True)

    def set_training_pairs(self, users, items, n_users, n_items, lists=False) -> None:
        """Positives as flat (user row, item row) arrays in file order; ``lists`` also builds the reference's
        dict-of-lists attributes (skipped for synthetic data at scale)."""
        users, items = np.asarray(users, np.int64), np.asarray(items, np.int64)
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.n_ratings = self.n_users * self.n_items
        by_u = np.argsort(users, kind='stable')
        by_i = np.argsort(items, kind='stable')
        u_ptr = np.zeros(self.n_users + 1, np.int64); np.cumsum(np.bincount(users, minlength=self.n_users), out=u_ptr[1:])
        i_ptr = np.zeros(self.n_items + 1, np.int64); np.cumsum(np.bincount(items, minlength=self.n_items), out=i_ptr[1:])
        self._csr = (u_ptr, items[by_u].astype(np.int32), i_ptr, users[by_i].astype(np.int32))
        if lists:
            self.usm = _lists(u_ptr, self._csr[1], self.n_users)
            self.ism = _lists(i_ptr, self._csr[3], self.n_items)
        self.u_rated = np.flatnonzero(np.diff(u_ptr) > 0).tolist()
        self.i_rated = np.flatnonzero(np.diff(i_ptr) > 0).tolist()
        self.fue = np.random.rand(self.n_users, self.k).astype(np.float32)
        self.fie = np.random.rand(self.n_items, self.k).astype(np.float32)
        self._sides = None

    def build_graph(self) -> None:
        tprint('%s does not require build_graph method!' % self.__sn)

    # ----------------------------------------------------------------- engine
    def _engine(self):
        if self._sides is None:
            u_ptr, u_idx, i_ptr, i_idx = self._csr
            self._sides = (topkrec.AlsSide(u_ptr, u_idx, self.seg, self.device), topkrec.AlsSide(i_ptr, i_idx, self.seg, self.device))
        return self._sides

    def _user_step(self, U, V):
        """``cer.py:36-46`` == ``wmf.py:67-77``."""
        us, it = self._engine()
        XX = topkrec.als_gram(V, it.rated_dev, self.b, self.lu)
        return topkrec.als_solve_rows(us, V, U, XX, self.a, self.b, 0.0, self.lu)

    def _item_step(self, U, V, prior=None):
        """``cer.py:47-63`` (prior given) / ``wmf.py:78-96``."""
        us, it = self._engine()
        XX = topkrec.als_gram(U, us.rated_dev, self.b, 0.0)
        return topkrec.als_solve_rows(it, U, V, XX, self.a, self.b, self.lv, self.lv, prior=prior,
                                      solve_empty=prior is not None, item_loss=True)

    def train(self, max_iter: int = 200, tol: float = 1e-4, model_path: str = None) -> None:
        loss = np.exp(50)
        if model_path is not None and os.path.isdir(model_path):
            self.import_embeddings(model_path)
        dev = torch.device(self.device)
        U = torch.from_numpy(np.ascontiguousarray(self.fue, np.float32)).to(dev)
        V = torch.from_numpy(np.ascontiguousarray(self.fie, np.float32)).to(dev)
        self.losses = []
        for it in range(max_iter):
            t1 = time.time()
            loss_old = loss
            loss = float(self._user_step(U, V).sum()) + float(self._item_step(U, V).sum())
            self.losses.append(loss)
            cond = np.abs(loss_old - loss) / loss_old
            tprint('Iter %3d, loss %.6f, converge %.6f, time %.2fs' % (it, loss, cond, time.time() - t1))
            if cond < tol:
                break
        self.fue, self.fie = U.cpu().numpy(), V.cpu().numpy()

    def export_model(self, model_path: str) -> None:
        return

    def import_model(self, model_path: str) -> None:
        return

"""Collaborative embedding regression on the B200 engine, behind the reference's ``CER`` surface
(``single/cer.py`` of domainxz/top-k-rec): WMF's alternation with a content prior ``V_j ~ F_j E`` on the items and a
ridge regression for ``E`` (``cer.py:24-73``).  The two half-steps are ``tkr_als_solve_rows`` calls; the one
``d_feat x d_feat`` system for ``E`` (``cer.py:27,64``) is factored once (fp64 Cholesky, a dense library call)."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from utils import tprint, get_embed_from_file, export_embed_to_file

from .wmf import WMF


class CER(WMF):
    def __init__(self, k: int, d: int, lu: float = 0.01, lv: float = 10, le: float = 10e3, a: float = 1, b: float = 0.01,
                 device: str = 'cuda', seg: int = 1024) -> None:
        super().__init__(k, lu, lv, a, b, device, seg)
        self.__sn = 'cer'
        self.d = d
        self.le = le
        self.E = None

    def train(self, max_iter: int = 200, tol: float = 1e-4, model_path: str = None) -> None:
        loss = np.exp(50)
        dev = torch.device(self.device)
        if model_path is not None and os.path.isdir(model_path):
            self.import_embeddings(model_path)
        if self.E is None:
            self.E = np.random.randn(self.feat.shape[1], self.k).astype(np.float32)
        F = torch.from_numpy(np.ascontiguousarray(self.feat, np.float32)).to(dev)
        F64 = F.double()
        FF = self.lv * (F64.T @ F64) + self.le * torch.eye(F.shape[1], dtype=torch.float64, device=dev)     # cer.py:27
        FF_chol = torch.linalg.cholesky(FF)
        del FF
        E = torch.from_numpy(np.ascontiguousarray(self.E, np.float64)).to(dev)
        U = torch.from_numpy(np.ascontiguousarray(self.fue, np.float32)).to(dev)
        V = torch.from_numpy(np.ascontiguousarray(self.fie, np.float32)).to(dev)
        self.losses = []
        for it in range(max_iter):
            t1 = time.time()
            Fe = (F64 @ E).float().contiguous()                                                             # cer.py:33
            loss_old = loss
            lu_rows = self._user_step(U, V)
            li_rows = self._item_step(U, V, prior=Fe)
            E = torch.cholesky_solve((self.lv * (F.T @ V)).double(), FF_chol)                               # cer.py:64
            loss = float(lu_rows.sum()) + float(li_rows.sum()) + 0.5 * self.le * float((E ** 2).sum())
            self.losses.append(loss)
            cond = np.abs(loss_old - loss) / loss_old
            tprint('Iter %3d, loss %.6f, time %.2fs' % (it, loss, time.time() - t1))
            if cond < tol:
                break
        Fe = (F64 @ E).float()
        unrated = torch.ones(self.n_items, dtype=torch.bool, device=dev)
        unrated[self._engine()[1].rated_dev.long()] = False
        V[unrated] = Fe[unrated]                                                                            # cer.py:70-73
        self.fue, self.fie, self.E = U.cpu().numpy(), V.cpu().numpy(), E.cpu().numpy()

    _E_FILE = 'final-E.dat'          # the content projection next to final-U/V.dat (cer.py:75-85)

    def import_model(self, model_path: str) -> None:
        path = os.path.join(model_path, self._E_FILE)
        if not os.path.exists(path):
            return
        tprint('Loading content projection matrix from %s' % path)
        self.E = get_embed_from_file(path)

    def export_model(self, model_path: str) -> None:
        if not os.path.exists(model_path) or getattr(self, 'E', None) is None:
            return
        path = os.path.join(model_path, self._E_FILE)
        tprint('Saving content projection matrix to %s' % path)
        export_embed_to_file(path, self.E)

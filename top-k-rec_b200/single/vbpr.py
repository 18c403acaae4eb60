"""VBPR (content-aware BPR) on the B200 engine behind the reference's ``VBPR`` class
(``single/vbpr.py`` of domainxz/top-k-rec).

The reference splits k into k/2 rating and k/2 content dimensions and feeds the dense
feature rows of the sampled items from the host every step (``vbpr.py:114``).  Here the
feature matrix lives on the device and the model is held directly in its export layout
(``vbpr.py:124-126``): ``U = [ur|uc]``, ``V = [ir | feat.cem]``, ``b = irb + feat.icb``
-- one ``tkr_vbpr_step`` per step (projection GEMM, fused BPR gather/scatter step, dE/dc
GEMM over the touched items, sparse + dense RMSProp).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

import topkrec
from utils import tprint

from .bpr import BPR


class VBPR(BPR):
    def __init__(self, k: int, d: int, lambda_u: float = 2.5e-3, lambda_i: float = 2.5e-3, lambda_j: float = 2.5e-4,
                 lambda_b: float = 0, lambda_e: float = 0, lr: float = 1.0e-4, mode: str = 'l2', graph: str = 'reference',
                 **engine_kw) -> None:
        """``graph='reference'`` (default): the objective exactly as ``vbpr.py:59-72`` computes it -- the bias variables are
        ``[n_items, 1]`` / ``[d, 1]``, so ``x_uij`` broadcasts to ``[B, B]`` (x[a, b] = r_a + y_b) and the loss sums over
        all B*B entries (a defect of the reference; O(B^2), batch_size <= 4096).
        ``graph='per-triple'``: x_n = r_n + y_n, the objective the code evidently means; any batch size."""
        super().__init__(k, lambda_u, lambda_i, lambda_j, lambda_b, lr, mode, **engine_kw)
        assert graph in ('reference', 'per-triple')
        self.graph = graph
        self.d = d
        self.le = lambda_e
        self.feat = None
        self._F = None

    def _engine_cfg(self):
        return topkrec.VbprCfg(self.n_users, self.n_items, self.k, self.d, self.lu, self.li, self.lj, self.lb, self.le,
                               self.lr, self.mode, self.optimizer, pairwise=self.graph == 'reference')

    def build_graph(self):
        """Device state per ``vbpr.py:37-48``: ur, uc, ir ~ N(0, 0.01); irb = 0; cem = 2/(d k); icb = 0."""
        assert self.k % 2 == 0, 'VBPR needs an even k'
        assert self.feat is not None, 'call load_content_data first'
        dev = torch.device(self.device)
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(self.seed) if self.seed is not None else int.from_bytes(os.urandom(7), 'little'))
        f32 = dict(dtype=torch.float32, device=dev)
        h = self.k // 2
        st = {'U': torch.randn(self.n_users, self.k, generator=gen, **f32) * 0.01,
              'V': torch.zeros(self.n_items, self.k, **f32),
              'rb': torch.zeros(self.n_items, **f32), 'bsum': torch.zeros(self.n_items, **f32),
              'E': torch.full((self.d, h), 2.0 / (self.d * self.k), **f32), 'c': torch.zeros(self.d, **f32)}
        st['V'][:, :h] = torch.randn(self.n_items, h, generator=gen, **f32) * 0.01
        for name in ('U', 'V', 'rb', 'E', 'c'):
            st['ms' + name] = torch.ones_like(st[name])
        self._state = st
        self._cfg = self._engine_cfg()
        self._ws = self._ws_batch = None
        if self._F is None or self._F.device != dev:
            self._F = torch.from_numpy(np.ascontiguousarray(self.feat, np.float32)).to(dev)
        return st

    def train(self, sampling: str = 'user uniform', epochs: int = 5, batch_size: int = 256,
              epoch_sample_limit: int = None, model_path: str = None) -> None:
        assert isinstance(sampling, str)
        assert isinstance(epochs, int)
        assert isinstance(batch_size, int)
        assert sampling == 'user uniform'
        if epoch_sample_limit is not None:
            self.epoch_sample_limit = int(epoch_sample_limit)          # vbpr.py:83-84 has no int assert
        batch_limit = self.epoch_sample_limit // batch_size + 1
        assert self.graph == 'per-triple' or batch_size <= 4096, \
            "graph='reference' reproduces the reference's [B, B] objective (O(B^2) per step): use batch_size <= 4096 or graph='per-triple'"
        self.build_graph()
        st, h = self._state, self.k // 2
        if model_path is not None:
            assert isinstance(model_path, str)
            tprint('Initialize weights with the previous trained model')
            self.import_embeddings(model_path)
        tprint('Training parameters: lu=%.6f, li=%.6f, lj=%.6f, lb=%.6f' % (self.lu, self.li, self.lj, self.lb))
        tprint('Learning rate is %.6f, regularization mode is %s' % (self.lr, self.mode))
        tprint('Training for %d epochs of %d batches using %s sampler' % (epochs, batch_limit, sampling))
        if self.fue is not None:
            tprint('Initialize user embeddings')
            self._assign('U', self.fue)
        if self.fie is not None:
            tprint('Initialize item embeddings')
            st['V'][:, :h] = torch.from_numpy(np.ascontiguousarray(self.fie[:, :h], np.float32)).to(st['V'].device)
        if self.fib is not None:
            tprint('Initialize item biases')
            # the reference assigns the EXPORTED bias (irb + feat.icb) back into irb (vbpr.py:106-108, SURVEY D-8)
            self._assign('rb', np.asarray(self.fib).ravel())
        topkrec.vbpr_project(self._cfg, st, self._F)
        self.losses = []
        for eid in range(epochs):
            t0 = time.time()
            self._run_steps(batch_limit - 1, batch_size, eid)
            sys.stderr.write(' ... total time collapse %10.4fs' % (time.time() - t0))
            sys.stderr.flush()
            print()
        self.fue = st['U'].cpu().numpy()
        self.fie = st['V'].cpu().numpy()                               # [ire | feat.cem]
        self.fib = st['bsum'].cpu().numpy().reshape(-1, 1)             # irb + feat.icb

    def _run_steps(self, n_steps, batch_size, eid=0):
        st = self._state
        if self._ws is None or self._ws_batch != batch_size:
            self._ws = topkrec.vbpr_workspace(self._cfg, batch_size, self.device)
            self._ws_batch = batch_size
            if self.tr_data:      # the most liked items: their gradient sums are privatised per thread block
                pos = np.fromiter((i for items in self.tr_data.values() for i in items), np.int64)
                topkrec.vbpr_set_hot_items(self._cfg, batch_size, self._ws, topkrec.popular_items(pos, self.n_items))
        chunk = max(1, min(n_steps, (1 << 20) // batch_size, 1024))
        host_gen = self._uniform_user_sampling(batch_size) if self.sampler_backend == 'numpy' else None
        done = 0
        while done < n_steps:
            n = min(chunk, n_steps - done)
            t1 = time.time()
            loss = torch.empty(n, dtype=torch.float32, device=st['U'].device)
            if host_gen is None:
                topkrec.vbpr_step(self._cfg, st, self._F, None, None, None, batch_size, n, self._ws, loss,
                                  sampler=self._device_sampler(), first_draw=self._draws)
                self._draws += n * batch_size
            else:
                trip = [next(host_gen) for _ in range(n)]
                u, i, j = (torch.from_numpy(np.concatenate([t[c] for t in trip])).to(st['U'].device) for c in range(3))
                topkrec.vbpr_step(self._cfg, st, self._F, u, i, j, batch_size, n, self._ws, loss)
            loss = loss.cpu().numpy()
            done += n
            self.losses.extend(loss.tolist())
            sys.stderr.write('\rEpoch=%3d, batch=%6d, loss=%8.2f, time=%4.4fs' % (eid + 1, done, loss[-1], (time.time() - t1) / n))
        return done

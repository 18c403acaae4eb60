"""VBPR (content-aware BPR) behind the reference's class surface (``single/vbpr.py``)."""
from __future__ import annotations

from .bpr import BPR


class VBPR(BPR):
    def __init__(self, k: int, d: int, lambda_u: float = 2.5e-3, lambda_i: float = 2.5e-3, lambda_j: float = 2.5e-4,
                 lambda_b: float = 0, lambda_e: float = 0, lr: float = 1.0e-4, mode: str = 'l2', **engine_kw) -> None:
        super().__init__(k, lambda_u, lambda_i, lambda_j, lambda_b, lr, mode, **engine_kw)
        self.d = d
        self.le = lambda_e
        self.feat = None

    def train(self, *args, **kwargs):
        raise NotImplementedError('VBPR step kernel (tkr_vbpr_step) is not built yet')

"""``DPM`` / ``ENCODER`` / ``MLP``: names kept so that ``from single import *`` (``single/__init__.py:1-9``, ``train.py:1``)
exposes what the reference exposes.  The model itself is dead code in the reference -- ``ENCODER`` declares the abstract
method ``pertrain`` (``encoder.py:22-24``) while ``MLP`` defines ``pretrain`` (``mlp.py:42``), so ``DPM.train`` raises
``TypeError`` at ``dpm.py:24`` before doing any work -- and is outside the hot-path scope (DPM row).
Here the failure is the same one line later in ``train.py:29-33``, with a message that says why."""
from __future__ import annotations

from abc import ABC, abstractmethod

from .wmf import WMF


class ENCODER(ABC):
    @abstractmethod
    def out(self):
        ...

    @abstractmethod
    def fit(self):
        ...

    @abstractmethod
    def pertrain(self, X, Y):          # sic (encoder.py:23)
        ...


class MLP(ENCODER):
    def __init__(self, *args, **kwargs) -> None:
        raise TypeError("MLP cannot be instantiated in the reference either (abstract 'pertrain', encoder.py:22-24 vs mlp.py:42); "
                        "the TensorFlow content encoder is outside this engine's scope")


class DPM(WMF):
    def __init__(self, k: int, d: int, lu: float = 0.01, lv: float = 10, le: float = 10e3, a: float = 1, b: float = 0.01, **kw) -> None:
        super().__init__(k, lu, lv, a, b, **kw)
        self.d = d
        self.le = le
        self.encoder = None

    def train(self, encoder=None, max_iter: int = 200, model_path: str = None) -> None:
        raise TypeError("DPM.train needs the reference's MLP encoder, which cannot be instantiated (encoder.py:22-24 vs mlp.py:42); "
                        "use CER for the content-regularised factorisation (single/cer.py)")

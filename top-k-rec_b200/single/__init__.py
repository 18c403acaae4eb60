"""Drop-in ``single`` package: the reference's model classes (``single/__init__.py:1-9``)
on the B200 engine.  ``from single import *`` + ``train.py:3-16`` run unchanged."""
from .rec import REC
from .bpr import BPR
from .vbpr import VBPR
from .wmf import WMF
from .cer import CER
from .dpm import DPM, ENCODER, MLP

__all__ = ['REC', 'BPR', 'VBPR', 'WMF', 'DPM', 'CER', 'ENCODER', 'MLP']

"""``REC``: the model interface of the reference (``single/rec.py:18-82``) --
same method names, attributes and on-disk ``final-{U,V,B}.dat`` contract."""
from __future__ import annotations

import os
import pickle
from abc import ABC, abstractmethod

import numpy as np
import scipy.sparse as ss

from utils import get_id_dict_from_file, tprint, export_embed_to_file, get_embed_from_file

_EMBED_FILES = (('fue', 'final-U.dat', 'user embeddings', 'uids'),
                ('fie', 'final-V.dat', 'item embeddings', 'iids'),
                ('fib', 'final-B.dat', 'item biases', 'iids'))


class REC(ABC):
    @abstractmethod
    def load_training_data(self):
        ...

    def load_content_data(self, content_file: str, iid_file: str) -> None:
        """Dense ``self.feat[n_items, d]`` aligned to ``self.iids`` from a pickled
        ndarray / scipy sparse matrix whose rows follow ``iid_file`` (``rec.py:23-33``)."""
        tprint('Load content data from %s' % content_file)
        file_rows = get_id_dict_from_file(iid_file)
        with open(content_file, 'rb') as f:
            raw = pickle.load(f, encoding='latin1')
        if ss.issparse(raw):
            raw = raw.tocsr()
        common = [iid for iid in self.iids if iid in file_rows]
        dst = np.fromiter((self.iids[i] for i in common), np.int64, count=len(common))
        src = np.fromiter((file_rows[i] for i in common), np.int64, count=len(common))
        self.feat = np.zeros((self.n_items, self.d), dtype=np.float32)
        block = raw[src]
        self.feat[dst] = block.toarray() if ss.issparse(block) else block
        tprint('Loading finished!')

    @abstractmethod
    def build_graph(self):
        ...

    @abstractmethod
    def train(self):
        ...

    @abstractmethod
    def export_model(self, model_path: str) -> None:
        ...

    def export_embeddings(self, model_path: str) -> None:
        """Write final-U/V/B.dat (whichever of fue/fie/fib exist) then the
        model-specific checkpoint (``rec.py:47-63``)."""
        if not os.path.exists(model_path):
            tprint('%s does not exist, create it instead' % model_path)
            os.makedirs(model_path)          # superset of the reference's non-recursive mkdir (D-6)
        if not os.path.isdir(model_path):
            tprint('%s is not a folder' % model_path)
            return
        for attr, fname, what, _ in _EMBED_FILES:
            if hasattr(self, attr):
                path = os.path.join(model_path, fname)
                tprint('Saving %s to %s' % (what, path))
                export_embed_to_file(path, getattr(self, attr))
        self.export_model(model_path)

    @abstractmethod
    def import_model(self, model_path: str) -> None:
        ...

    def import_embeddings(self, model_path: str) -> None:
        """Read back whichever final-*.dat exist, then the checkpoint (``rec.py:69-82``)."""
        for attr, fname, what, ids in _EMBED_FILES:
            path = os.path.join(model_path, fname)
            if os.path.exists(path):
                tprint('Loading %s from %s' % (what, path))
                setattr(self, attr, get_embed_from_file(path, getattr(self, ids)))
        self.import_model(model_path)

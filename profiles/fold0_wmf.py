#!/usr/bin/env python3
"""train.py:18-22 (the WMF block: WMF(k=50), train(max_iter=200, tol=1e-4), export, warm restart with max_iter=20) on the B200
engine and the shipped fold 0; the first iteration is compared with oracle/als_ref.py run from the same seeded start
(profiles/fold0_wmf_oracle.py).  Reads only <data_dir>.
usage: python profiles/fold0_wmf.py <data_dir>"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "top-k-rec_b200"))
from single import WMF  # noqa: E402
import evaluate  # noqa: E402


def rel(x, ref):
    return float(np.abs(np.asarray(x, np.float64) - ref).max() / np.abs(ref).max())


D = sys.argv[1]
ref = np.load(os.path.join(D, "wmf_oracle.npz"))
model = WMF(k=50)
np.random.seed(2027)
model.load_training_data(D + "/uid", D + "/vid", D + "/f0tr.txt")
start = (model.fue.copy(), model.fie.copy())
with contextlib.redirect_stdout(io.StringIO()):
    model.train(max_iter=1, tol=0.0)
out = {"workload": "train.py:18-22 WMF(k=50) on the shipped fold 0", "first_iteration_rel_diff_U": rel(model.fue, ref["fue"].astype(np.float64)),
       "first_iteration_rel_diff_V": rel(model.fie, ref["fie"].astype(np.float64)), "first_iteration_loss": model.losses[0],
       "oracle_first_iteration_loss": float(ref["losses"][0]), "oracle_seconds_per_iteration": float(ref["seconds"]), "oracle_cores": int(ref["cores"])}
model.fue, model.fie = start
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "wmf")
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        model.train(max_iter=200, tol=1e-4)                                # train.py:20
    out["train_seconds"] = time.time() - t0
    out["iterations_to_tol_1e-4"] = len(model.losses); out["losses_first_last"] = [model.losses[0], model.losses[-1]]
    with contextlib.redirect_stdout(io.StringIO()):
        model.export_embeddings(path)                                      # train.py:21
        model.train(max_iter=20, model_path=path)                          # train.py:22 (warm restart from the .dat files)
        out["warm_restart_iterations"] = len(model.losses)
        model.export_embeddings(path)
        out["accuracy_im"] = evaluate.main(["-d", D, "-m", path, "-f", "0", "-sl", "im"])[0]
print(json.dumps(out))

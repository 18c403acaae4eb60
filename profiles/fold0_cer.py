#!/usr/bin/env python3
"""train.py:24-27 (the CER block: CER(k=50, d=20000) on the shipped fold 0 with data/meta.pkl) on the B200 engine, from the
same seeded start as profiles/fold0_cer_reference.py (the unmodified reference, run in the build container), then both
models through the accelerated evaluator.  Reads only <data_dir> (a copy of the reference's data/ files + cer_reference.npz).
usage: python profiles/fold0_cer.py <data_dir> [iters=2]"""
import contextlib
import io
import json
import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "top-k-rec_b200")
sys.path.insert(0, PKG)
from single import CER  # noqa: E402
import evaluate  # noqa: E402


def rel(x, ref):
    return float(np.abs(np.asarray(x, np.float64) - ref).max() / np.abs(ref).max())


def main():
    D = sys.argv[1]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ref = np.load(os.path.join(D, "cer_reference.npz"))
    m = CER(k=50, d=20000)
    np.random.seed(2026)
    t0 = time.time()
    m.load_training_data(D + "/uid", D + "/vid", D + "/f0tr.txt")
    m.load_content_data(D + "/meta.pkl", D + "/vid")
    m.E = np.random.randn(m.feat.shape[1], m.k).astype(np.float32)
    t_load = time.time() - t0
    buf = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(buf):
        m.train(max_iter=iters, tol=0.0)
    t_train = time.time() - t0
    it_s = [float(x) for x in re.findall(r"time ([0-9.]+)s", buf.getvalue())]
    out = {"workload": "train.py:24-27 CER(k=50, d=20000) on the shipped fold 0 (69 878 users, 10 380 items, 919 952 positives, meta.pkl)",
           "iterations": iters, "losses": m.losses, "reference_losses": ref["losses"].tolist(),
           "loss_rel_diff": [abs(a - b) / abs(b) for a, b in zip(m.losses, ref["losses"])],
           "rel_diff_U": rel(m.fue, ref["fue"].astype(np.float64)), "rel_diff_V": rel(m.fie, ref["fie"].astype(np.float64)),
           "rel_diff_E": rel(m.E, ref["E"].astype(np.float64)),
           "load_seconds": t_load, "train_seconds_total": t_train, "seconds_per_iteration": it_s,
           "reference_seconds_per_iteration": float(ref["seconds"]) / int(ref["iters"]), "reference_cores": int(ref["cores"])}
    with tempfile.TemporaryDirectory() as td:
        ours, theirs = os.path.join(td, "ours"), os.path.join(td, "ref")
        m.export_embeddings(ours)
        r = CER(k=50, d=20000); r.uids, r.iids = m.uids, m.iids
        r.fue, r.fie, r.E = ref["fue"], ref["fie"], ref["E"]
        r.export_embeddings(theirs)
        with contextlib.redirect_stdout(io.StringIO()):
            out["accuracy_im_ours"] = evaluate.main(["-d", D, "-m", ours, "-f", "0", "-sl", "im"])[0]
            out["accuracy_im_reference_model"] = evaluate.main(["-d", D, "-m", theirs, "-f", "0", "-sl", "im"])[0]
        out["final_E_dat_written"] = os.path.exists(os.path.join(ours, "final-E.dat"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

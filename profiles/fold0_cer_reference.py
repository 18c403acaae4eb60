#!/usr/bin/env python3
"""Build-container only: run the UNMODIFIED reference CER (single/cer.py via oracle/tf_stub) on the shipped fold 0
(train.py:24-27: CER(k=50, d=20000), data/meta.pkl) for a few iterations from a seeded start and save its factors and printed
losses, for profiles/fold0_cer.py to compare the B200 engine against on the GPU box.
usage: python profiles/fold0_cer_reference.py <out_dir> [iters=2]"""
import contextlib
import io
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import import_reference  # noqa: E402

out, iters = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
single, _ = import_reference()
D = "/root/reference/data"
m = single.CER(k=50, d=20000)
np.random.seed(2026)
m.load_training_data(D + "/uid", D + "/vid", D + "/f0tr.txt")
m.load_content_data(D + "/meta.pkl", D + "/vid")
m.E = np.random.randn(m.feat.shape[1], m.k).astype(np.float32)
buf = io.StringIO()
t0 = time.time()
with contextlib.redirect_stdout(buf):
    m.train(max_iter=iters, tol=0.0)
dt = time.time() - t0
losses = [float(x) for x in re.findall(r"loss ([-0-9.e+]+),", buf.getvalue())]
np.savez(os.path.join(out, "cer_reference.npz"), fue=m.fue, fie=m.fie, E=m.E, losses=np.array(losses), seconds=dt, iters=iters, cores=os.cpu_count())
print("reference CER on fold 0: %d iterations in %.1f s, losses %s" % (iters, dt, losses))

#!/usr/bin/env python3
"""Build-container only: ONE iteration of the intended WMF alternation (oracle/als_ref.py; the reference's shipped
WMF.train raises at wmf.py:72) on the shipped fold 0 from a seeded start, for profiles/fold0_wmf.py to compare against.
The start is drawn exactly as the reference's loader does (np.random.rand for fue then fie, wmf.py:55-56).
usage: python profiles/fold0_wmf_oracle.py <out_dir>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import import_reference  # noqa: E402
from oracle import als_ref  # noqa: E402

out = sys.argv[1]
single, _ = import_reference()
D = "/root/reference/data"
m = single.WMF(k=50)
np.random.seed(2027)
m.load_training_data(D + "/uid", D + "/vid", D + "/f0tr.txt")          # the reference's own loader + start
u_ptr, u_idx = als_ref.csr_from_lists(m.usm, m.n_users)
i_ptr, i_idx = als_ref.csr_from_lists(m.ism, m.n_items)
t0 = time.time()
U, V, losses = als_ref.wmf_train(m.fue, m.fie, u_ptr, u_idx, i_ptr, i_idx, m.a, m.b, m.lu, m.lv, max_iter=1, tol=0.0)
np.savez(os.path.join(out, "wmf_oracle.npz"), fue=U, fie=V, losses=np.array(losses), seconds=time.time() - t0, cores=os.cpu_count())
print("oracle WMF on fold 0: 1 iteration in %.1f s, loss %s" % (time.time() - t0, losses))

#!/usr/bin/env python3
"""Golden vectors for the ALS path (SURVEY.md 8(f) NEXT-1), produced by the UNMODIFIED reference.

Run ONLY in the build container (needs /root/reference, read-only):

    python tests/golden/make_golden_als.py

``als_cer.npz``: ``CER.load_training_data`` + ``CER.train`` of the REFERENCE (single/wmf.py:33-56,
single/cer.py:24-73, imported through oracle/tf_stub) on tests/golden/mini with a seeded random start
(``np.random.seed(77)`` drives the reference's own ``np.random.rand`` / ``randn`` initialisers, wmf.py:55-56,
cer.py:31) and a seeded dense content matrix: the start state, the state after ``max_iter`` = 1 and 4 iterations,
and the losses the reference printed.
"""
import contextlib
import io
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def run(single, mini, k, d_feat, iters, tol):
    m = single.CER(k=k, d=d_feat)
    np.random.seed(77)
    with tempfile.TemporaryDirectory() as td:
        # the WMF loader raises KeyError on ids missing from uid/vid (wmf.py:51); mini's last line is such a user
        tr = os.path.join(td, "tr.txt")
        open(tr, "w").write("".join(ln for ln in open(os.path.join(mini, "f0tr.txt")) if not ln.startswith("99999,")))
        m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), tr)
    rng = np.random.default_rng(78)
    m.feat = (np.abs(rng.standard_normal((m.n_items, d_feat))) * (rng.random((m.n_items, d_feat)) < 0.4)).astype(np.float32)
    m.E = np.random.randn(d_feat, k).astype(np.float32)          # what cer.py:31 would draw
    start = (m.fue.copy(), m.fie.copy(), m.E.copy(), m.feat.copy())
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.train(max_iter=iters, tol=tol)
    losses = [float(x) for x in re.findall(r"loss ([-0-9.e+]+),", buf.getvalue())]
    return m, start, losses


def main():
    single, _ = import_reference()
    mini = os.path.join(HERE, "mini")
    k, d_feat = 12, 9
    m1, start, l1 = run(single, mini, k, d_feat, 1, 0.0)
    m4, start4, l4 = run(single, mini, k, d_feat, 4, 0.0)
    assert all(np.array_equal(x, y) for x, y in zip(start, start4)) and abs(l1[0] - l4[0]) < 1e-6 * abs(l1[0])
    u_rows = np.array([len(m1.usm[u]) for u in range(m1.n_users)])
    i_rows = np.array([len(m1.ism[j]) for j in range(m1.n_items)])
    assert (u_rows == 0).any() and (i_rows == 0).any()            # both kinds of empty rows are covered
    np.savez_compressed(
        os.path.join(HERE, "als_cer.npz"), k=k, d_feat=d_feat, a=m1.a, b=m1.b, lu=m1.lu, lv=m1.lv, le=m1.le,
        fue0=start[0], fie0=start[1], E0=start[2], feat=start[3],
        u_idx=np.concatenate([np.asarray(m1.usm[u], np.int32) for u in range(m1.n_users)]), u_cnt=u_rows,
        i_idx=np.concatenate([np.asarray(m1.ism[j], np.int32) for j in range(m1.n_items)]), i_cnt=i_rows,
        fue1=m1.fue, fie1=m1.fie, E1=m1.E, fue4=m4.fue, fie4=m4.fie, E4=m4.E, losses=np.array(l4))
    print("als_cer.npz: users %d (%d empty), items %d (%d empty), nnz %d, losses %s" % (
        m1.n_users, (u_rows == 0).sum(), m1.n_items, (i_rows == 0).sum(), u_rows.sum(), l4))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""train.py:3-16 (the BPR and VBPR blocks, verbatim calls) on the shipped fold 0 on the B200 engine, then the GPU
evaluator against the CPU oracle on the trained models: per scenario the filtered top-30 lists of ALL 69 878 users
(GPU, tcgen05 engine) must equal oracle/evaluate_ref.py's (exact fp32 FMA chains; pinned to the unmodified evaluate.py
by tests/golden), with and without final-B.dat.  The models go to <out_dir> as .dat (gzipped tar) so that the unmodified
reference evaluate.py can be run on the same files in the build container (profiles/fold0_bpr_check.py).
Reads only <data_dir> (a copy of the reference's data/ files).
usage: python profiles/fold0_bpr.py <data_dir> <out_dir>"""
import contextlib
import io
import json
import os
import shutil
import sys
import tarfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "top-k-rec_b200")
sys.path[:0] = [PKG, ROOT]
from single import BPR, VBPR  # noqa: E402
import evaluate  # noqa: E402
from oracle import evaluate_ref  # noqa: E402  (the checker)


def crc_rows(lists):
    """per-user CRC32 of the 30 int32 columns; one CRC over those = the fingerprint of a whole scenario"""
    per = np.fromiter((zlib.crc32(r.tobytes()) for r in np.ascontiguousarray(lists, np.int32)), np.uint32, count=lists.shape[0])
    return per, int(zlib.crc32(per.tobytes()))


def train_block(model, D, out, content=False):
    """the five calls of a train.py block; stdout of the engine's progress lines is swallowed"""
    t = {}
    t0 = time.time()
    model.load_training_data(D + "/uid", D + "/vid", D + "/f0tr.txt")
    if content:
        model.load_content_data(D + "/meta.pkl", D + "/vid")
    t["load_s"] = time.time() - t0
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        model.train(epochs=5, batch_size=256, epoch_sample_limit=10e5)
    t["train_s"] = time.time() - t0
    first = list(model.losses)
    model.export_embeddings(out)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        model.train(epochs=5, batch_size=256, epoch_sample_limit=10e5, model_path=out)
    t["warm_train_s"] = time.time() - t0
    model.export_embeddings(out)
    steps = len(first)
    t.update(steps_per_train=steps, triples_per_train=steps * 256, loss_first=first[0], loss_last=first[-1],
             loss_last_warm=model.losses[-1], train_triples_per_s=steps * 256 / t["train_s"])
    return t


def check_model(D, mdir, name):
    res = {}
    for use_bias in (True, False):
        tag = "with_bias" if use_bias else "no_bias"
        t0 = time.time()
        lines, lists = evaluate.run(D, mdir, 0, 5, 30, ("im", "om"), use_bias=use_bias, keep_lists=True)
        t_gpu = time.time() - t0
        t0 = time.time()
        ref = evaluate_ref.evaluate(D, mdir, scenarios=("im", "om"), use_bias=use_bias)
        t_cpu = time.time() - t0
        r = {"gpu_lines": lines, "oracle_lines": [evaluate_ref.format_line(s, ref[s][0]) for s in ("im", "om")],
             "gpu_evaluate_s": t_gpu, "oracle_evaluate_s": t_cpu}
        for sc in ("im", "om"):
            g, o = lists[sc], ref[sc][1]
            per_g, crc_g = crc_rows(g)
            per_o, crc_o = crc_rows(o)
            r[sc] = {"users": int(g.shape[0]), "users_with_different_list": int((per_g != per_o).sum()),
                     "elements_different": int((g != o).sum()), "crc_of_user_crcs_gpu": crc_g, "crc_of_user_crcs_oracle": crc_o}
        r["lines_equal"] = r["gpu_lines"] == r["oracle_lines"]
        res[tag] = r
        print(name, tag, lines, "oracle:", r["oracle_lines"], {sc: r[sc]["users_with_different_list"] for sc in ("im", "om")}, flush=True)
    return res


def main():
    D, OUT = sys.argv[1], sys.argv[2]
    os.makedirs(OUT, exist_ok=True)
    out = {"workload": "train.py:3-16 on the shipped fold 0 (69 878 users, 10 380 items, 919 952 positives): BPR(k=50) and "
                       "VBPR(k=50, d=20000, meta.pkl), epochs=5, batch_size=256, epoch_sample_limit=10e5, then a warm-start train"}
    np.random.seed(2026)
    for name, ctor, content in (("bpr", lambda: BPR(k=50, seed=2026), False), ("vbpr", lambda: VBPR(k=50, d=20000, seed=2026), True)):
        mdir = os.path.join(OUT, name)
        shutil.rmtree(mdir, ignore_errors=True)
        out[name] = {"train": train_block(ctor(), D, mdir, content)}
        print(name, json.dumps(out[name]["train"]), flush=True)
        out[name]["evaluate"] = check_model(D, mdir, name)
    with tarfile.open(os.path.join(OUT, "fold0_models.tar.gz"), "w:gz") as tf:
        for name in ("bpr", "vbpr"):
            for f in ("final-U.dat", "final-V.dat", "final-B.dat"):
                tf.add(os.path.join(OUT, name, f), arcname="%s/%s" % (name, f))
    for name in ("bpr", "vbpr"):
        shutil.rmtree(os.path.join(OUT, name), ignore_errors=True)
    json.dump(out, open(os.path.join(OUT, "fold0_bpr.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

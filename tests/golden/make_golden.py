#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

Run ONLY in the build container (needs /root/reference, read-only):

    python tests/golden/make_golden.py            # mini fixtures (seconds)
    python tests/golden/make_golden.py --fold0    # + shipped fold 0 numbers (~2 min)

What comes from where
---------------------
* ``mini/``                     a small synthetic data set in the reference's file
                                formats (README.md:56-69), generated here (seeded).
* ``loader.json``               ``BPR.load_training_data`` of the REFERENCE
                                (single/bpr.py:51-69) on mini/, imported through
                                oracle/tf_stub.
* ``sampler.npz``               first batches of the REFERENCE
                                ``BPR._uniform_user_sampling`` (single/bpr.py:155-165)
                                after ``np.random.seed(123)``.
* ``codec.dat`` / ``codec.npy`` REFERENCE ``utils.export_embed_to_file`` output and
                                ``utils.get_embed_from_file`` read-back (utils.py:28-55).
* ``mini_model*/final-*.dat``   models written by the REFERENCE ``REC.export_embeddings``.
* ``evaluate_mini.json``        stdout of the UNMODIFIED ``/root/reference/evaluate.py``
                                on mini/ (subprocess).
* ``evaluate_fold0.json``       same on the shipped fold 0 with a seeded random model
                                (d=50) + how many users' top-30 sets differ between the
                                BLAS route and the FMA-chain oracle (ties / near-ties).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def make_mini(out):
    rng = np.random.default_rng(20261017)
    os.makedirs(out, exist_ok=True)
    n_users, n_items, n_om = 300, 160, 40
    uid = [str(x) for x in rng.permutation(np.arange(1, 2000))[:n_users]]
    vid = [str(1000 + 3 * x) for x in range(n_items)]
    perm = rng.permutation(n_items)
    om = sorted(perm[:n_om].tolist()); im = sorted(perm[n_om:].tolist())
    open(os.path.join(out, "uid"), "w").write("".join(u + "\n" for u in uid))
    open(os.path.join(out, "vid"), "w").write("".join(v + "\n" for v in vid))
    for name, cols in (("f0tr.idl", im), ("f0te.im.idl", im), ("f0te.om.idl", om), ("f0te.all.idl", range(n_items))):
        open(os.path.join(out, name), "w").write("".join(vid[c] + "\n" for c in cols))
    pop = 1.0 / np.arange(1, len(im) + 1); pop /= pop.sum()
    tr, te_im, te_om, te_all = [], [], [], []
    for r in rng.permutation(n_users):
        n_r = int(rng.integers(5, 60))
        items = rng.choice(im, size=n_r, replace=False, p=pop)
        likes = rng.random(n_r) < (0.0 if r % 37 == 0 else 0.2)      # some users without positives
        tr.append(uid[r] + "".join(",%s:%d" % (vid[c], l) for c, l in zip(items, likes)))
        rest = np.setdiff1d(im, items)
        t_items = rng.choice(rest, size=min(len(rest), int(rng.integers(3, 25))), replace=False)
        if r % 11 == 0:                                                # a liked test item that is also rated
            t_items = np.concatenate([t_items, items[:2]])
        t_like = rng.random(len(t_items)) < 0.35
        te_im.append(uid[r] + "".join(",%s:%d" % (vid[c], l) for c, l in zip(t_items, t_like)))
        o_items = rng.choice(om, size=int(rng.integers(1, 12)), replace=False)
        o_like = rng.random(len(o_items)) < 0.35
        te_om.append(uid[r] + "".join(",%s:%d" % (vid[c], l) for c, l in zip(o_items, o_like)))
        te_all.append(te_im[-1] + "".join(",%s:%d" % (vid[c], l) for c, l in zip(o_items, o_like)))
    tr.append("99999,1000:1")                                          # user unknown to uid file (utils.py:63)
    for name, lines in (("f0tr.txt", tr), ("f0te.im.txt", te_im), ("f0te.om.txt", te_om), ("f0te.all.txt", te_all)):
        open(os.path.join(out, name), "w").write("".join(l + "\n" for l in lines))
    return n_users, n_items


def import_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_stub"))
    sys.path.insert(0, REF)
    import single  # noqa: F401  (the reference package)
    import utils as ref_utils
    assert os.path.abspath(single.__file__).startswith(REF) and os.path.abspath(ref_utils.__file__).startswith(REF)
    return single, ref_utils


def run_reference_evaluate(data, model, scenarios, fold=0):
    out = subprocess.check_output([sys.executable, os.path.join(REF, "evaluate.py"), "-d", data, "-m", model,
                                   "-f", str(fold), "-sl"] + list(scenarios), text=True)
    return [ln for ln in out.strip().splitlines() if ln]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fold0", action="store_true")
    args = ap.parse_args()
    single, ref_utils = import_reference()
    from oracle import evaluate_ref

    mini = os.path.join(HERE, "mini")
    n_users, n_items = make_mini(mini)

    # ---- loader + sampler (reference objects) ---------------------------------
    m = single.BPR(k=8)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    json.dump({"n_users": m.n_users, "n_items": m.n_items, "epoch_sample_limit": m.epoch_sample_limit,
               "tr_users": [int(u) for u in m.tr_users],
               "tr_data": {str(u): [int(x) for x in v] for u, v in m.tr_data.items()}},
              open(os.path.join(HERE, "loader.json"), "w"))
    np.random.seed(123)
    gen = m._uniform_user_sampling(64)
    ub, ib, jb = [], [], []
    for _ in range(12):
        a, b, c = next(gen)
        ub.append(np.array(a, np.int32)); ib.append(b.copy()); jb.append(c.copy())
    np.savez(os.path.join(HERE, "sampler.npz"), seed=123, batch=64, ub=np.array(ub), ib=np.array(ib), jb=np.array(jb))

    # ---- .dat codec -----------------------------------------------------------
    rng = np.random.default_rng(7)
    emb = np.concatenate([rng.standard_normal((5, 7)).astype(np.float32) * np.float32(0.01),
                          np.array([[0, -0.0, 1e-7, -1e-7, 123456.789, -5e-7, 4.9999995e-7]], np.float32)])
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x", "codec.dat"); os.mkdir(os.path.dirname(p))
        ref_utils.export_embed_to_file(p, emb)
        data = open(p, "rb").read()
        back = ref_utils.get_embed_from_file(p)
    open(os.path.join(HERE, "codec.dat"), "wb").write(data)
    np.save(os.path.join(HERE, "codec.npy"), np.stack([emb, back]))

    # ---- models written by the reference's export_embeddings -----------------
    k = 16
    rng = np.random.default_rng(11)
    U = (0.1 * rng.standard_normal((n_users, k))).astype(np.float32)
    V = (0.1 * rng.standard_normal((n_items, k))).astype(np.float32)
    Bv = (0.05 * rng.standard_normal((n_items, 1))).astype(np.float32)
    Ut, Vt = U.copy(), V.copy()
    Vt[5] = Vt[17]; Vt[40] = Vt[17]; Vt[60] = 0; Vt[61] = 0; Vt[100] = 0   # exact ties
    Ut[3] = 0                                                               # a user whose scores all tie
    for name, uu, vv, bb in (("mini_model", U, V, None), ("mini_model_bias", U, V, Bv), ("mini_model_ties", Ut, Vt, None)):
        mm = single.BPR(k=k); mm.uids, mm.iids = m.uids, m.iids
        mm.fue, mm.fie = uu, vv
        if bb is not None:
            mm.fib = bb
        else:
            del mm.fib                       # rec.py:58 uses hasattr -> no final-B.dat
        mm.export_embeddings(os.path.join(HERE, name))

    # ---- reference evaluate.py on mini ---------------------------------------
    scs = ("im", "om", "all")
    fmt = evaluate_ref.format_line
    res = {"no_bias": run_reference_evaluate(mini, os.path.join(HERE, "mini_model"), scs),
           # evaluate.py:80 only broadcasts when n_te == n_items -> scenario 'all' is the one bias case it can run
           "bias_all": run_reference_evaluate(mini, os.path.join(HERE, "mini_model_bias"), ["all"]),
           "ties_reference_unstable_sort": run_reference_evaluate(mini, os.path.join(HERE, "mini_model_ties"), scs)}
    ours = evaluate_ref.evaluate(mini, os.path.join(HERE, "mini_model"), scenarios=scs)
    ours_b = evaluate_ref.evaluate(mini, os.path.join(HERE, "mini_model_bias"), scenarios=("all",))
    ours_t = evaluate_ref.evaluate(mini, os.path.join(HERE, "mini_model_ties"), scenarios=scs)
    res["ties_oracle_stable"] = [fmt(sc, ours_t[sc][0]) for sc in scs]
    # the literal restatement (np.dot + default unstable argsort) must hit the reference even under ties
    import functools
    from oracle import topk_ref
    stable_fn = topk_ref.score_topk_numpy
    topk_ref.score_topk_numpy = functools.partial(stable_fn, stable=False)
    lit = evaluate_ref.evaluate(mini, os.path.join(HERE, "mini_model_ties"), scenarios=scs, scorer="blas")
    topk_ref.score_topk_numpy = stable_fn
    print(json.dumps(res, indent=1))
    assert res["no_bias"] == [fmt(sc, ours[sc][0]) for sc in scs], "oracle != reference evaluate.py on mini (no bias)"
    assert res["bias_all"] == [fmt("all", ours_b["all"][0])], "oracle != reference evaluate.py on mini (bias)"
    assert res["ties_reference_unstable_sort"] == [fmt(sc, lit[sc][0]) for sc in scs], "literal restatement != reference under ties"
    json.dump(res, open(os.path.join(HERE, "evaluate_mini.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "evaluate_mini_lists.npz"),
                        **{sc: ours[sc][1] for sc in ours}, all_bias=ours_b["all"][1],
                        **{"ties_" + sc: ours_t[sc][1] for sc in ours_t})

    if args.fold0:
        import time
        data = os.path.join(REF, "data")
        uids = ref_utils.get_id_dict_from_file(os.path.join(data, "uid"))
        vids = ref_utils.get_id_dict_from_file(os.path.join(data, "vid"))
        rng = np.random.default_rng(50)
        with tempfile.TemporaryDirectory() as td:
            md = os.path.join(td, "model"); os.mkdir(md)
            U0 = (0.1 * rng.standard_normal((len(uids), 50))).astype(np.float32)
            V0 = (0.1 * rng.standard_normal((len(vids), 50))).astype(np.float32)
            V0[100:110] = V0[99]; V0[500:520] = 0
            from oracle.codec_ref import dat_bytes
            open(os.path.join(md, "final-U.dat"), "wb").write(dat_bytes(U0))
            open(os.path.join(md, "final-V.dat"), "wb").write(dat_bytes(V0))
            t0 = time.time(); ref_lines = run_reference_evaluate(data, md, ["im", "om"]); t_ref = time.time() - t0
            t0 = time.time(); fma = evaluate_ref.evaluate(data, md, scenarios=("im", "om")); t_fma = time.time() - t0
            blas = evaluate_ref.evaluate(data, md, scenarios=("im", "om"), scorer="blas")
        rep = {"reference_stdout": ref_lines, "reference_wall_s": t_ref, "oracle_wall_s": t_fma,
               "oracle_fma": [evaluate_ref.format_line(sc, fma[sc][0]) for sc in ("im", "om")],
               "oracle_blas_stable": [evaluate_ref.format_line(sc, blas[sc][0]) for sc in ("im", "om")],
               "users_with_different_top30_set_fma_vs_blas": {
                   sc: int(sum(set(a) != set(b) for a, b in zip(fma[sc][1].tolist(), blas[sc][1].tolist()))) for sc in ("im", "om")},
               "model": "U,V ~ 0.1*N(0,1) default_rng(50), d=50, V[100:110]=V[99], V[500:520]=0, 6-decimal .dat text"}
        print(json.dumps(rep, indent=1))
        json.dump(rep, open(os.path.join(HERE, "evaluate_fold0.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

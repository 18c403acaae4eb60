"""CPU tests of the host side: file formats and loaders against the reference-generated
goldens, the class surface, the C ABI (symbols only -- no compute without a GPU), and the
"no fallback" rule."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import utils
import topkrec
import single
from conftest import ROOT


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "topkrec.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(tkr_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 12
    lib = ctypes.CDLL(os.path.join(ROOT, "top-k-rec_b200", "topkrec", "libtopkrec.so"))
    for name in declared:
        assert hasattr(lib, name), "libtopkrec.so does not export %s" % name
    assert topkrec.version() == int(re.search(r"#define TKR_VERSION (\d+)", header).group(1))


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable here."""
    L = topkrec.lib()
    assert L.tkr_bpr_workspace_bytes(None, 256) == 0
    cfg = topkrec.BprCfg(10, 10, 8)
    assert L.tkr_bpr_workspace_bytes(cfg.ptr, 256) > 0
    rc = L.tkr_bpr_step(cfg.ptr, None, None, None, None, None, None, None, None, None, 256, 1, None, 0, None, None, 0, None)
    assert rc == -1 and b"must not be NULL" in L.tkr_last_error()
    rc = L.tkr_score_topk(1, 4, 1, 4, 8, None, None, None, 65, 0, 1, 1, None, 0, None)
    assert rc == -1 and b"k must be in" in L.tkr_last_error()
    rc = L.tkr_topk_merge(None, None, 2, 4, 5, None, None, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    U = torch.zeros(4, 8); V = torch.zeros(6, 8)
    with pytest.raises(topkrec.TkrError):
        topkrec.score_topk(U, V, 3)
    cfg = topkrec.BprCfg(4, 6, 8)
    with pytest.raises((topkrec.TkrError, RuntimeError, AssertionError)):
        topkrec.bpr_workspace(cfg, 16, "cuda")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "top-k-rec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


# ------------------------------------------------------------------ formats vs reference goldens
def test_dat_writer_is_byte_identical(golden, tmp_path):
    emb, back = np.load(os.path.join(golden, "codec.npy"))
    p = tmp_path / "m" / "x.dat"
    utils.export_embed_to_file(str(p), emb)
    assert p.read_bytes() == open(os.path.join(golden, "codec.dat"), "rb").read()
    assert np.array_equal(utils.get_embed_from_file(str(p)), back)
    ids = {"a": 2, "b": 0, "c": 1}
    sub = utils.get_embed_from_file(str(p), ids)
    assert sub.shape == (3, emb.shape[1]) and np.array_equal(sub, back[:3])
    assert utils.get_embed_from_file(str(tmp_path / "missing.dat")) is None


def test_reference_written_model_reads_back(golden):
    U = utils.get_embed_from_file(os.path.join(golden, "mini_model_bias", "final-U.dat"))
    B = utils.get_embed_from_file(os.path.join(golden, "mini_model_bias", "final-B.dat"))
    assert U.shape == (300, 16) and B.shape == (160, 1) and U.dtype == np.float32


def test_loaders_match_reference(golden, mini):
    g = json.load(open(os.path.join(golden, "loader.json")))
    m = single.BPR(k=8)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"), data_copy=True)
    assert (m.n_users, m.n_items, m.epoch_sample_limit) == (g["n_users"], g["n_items"], g["epoch_sample_limit"])
    assert m.tr_users == g["tr_users"]
    assert {str(k): v for k, v in m.tr_data.items()} == g["tr_data"]
    assert len(m.data) == g["epoch_sample_limit"] and isinstance(m.data[0], tuple)
    assert utils.get_id_dict_from_file("/nonexistent") == {} and utils.get_data_from_file("/nonexistent", {}, {}) == []
    ivt = utils.get_iv_dict_from_file(os.path.join(mini, "vid"))
    assert ivt[0] == "1000" and len(ivt) == 160


def test_native_loader_matches_reference(golden, mini):
    """the tkr_ratings_parse route (data_copy=False) builds the same structures as the reference's loader"""
    g = json.load(open(os.path.join(golden, "loader.json")))
    m = single.BPR(k=8)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    assert (m.n_users, m.n_items, m.epoch_sample_limit) == (g["n_users"], g["n_items"], g["epoch_sample_limit"])
    assert m.tr_users == g["tr_users"]
    assert {str(k): v for k, v in m.tr_data.items()} == g["tr_data"]


def test_native_codec_edge_cases(tmp_path):
    """tkr_dat_write / tkr_dat_read against Python's own '%f' and float(): signed zeros, denormals, huge values,
    ragged whitespace, exponent notation (the legacy C++ writer's '%10.8e', old/cr/utils.cpp:90-97), empty matrix"""
    import topkrec
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((257, 19)) * rng.choice([1e-6, 1e-3, 1.0, 1e4, 1e9], (257, 19))).astype(np.float32)
    x[0, :10] = [0.0, -0.0, 1e-7, -1e-7, 123456.789, 3.4e38, -3.4e38, 1e-45, 0.5000005, 2.5000015]
    p = str(tmp_path / "x.dat")
    topkrec.dat_write(p, x)
    want = "".join("".join("%f " % v for v in row) + "\n" for row in x.astype(np.float64))
    assert open(p).read() == want
    back = topkrec.dat_read(p)
    ref = np.array([[np.float32(t) for t in line.split()] for line in want.splitlines()], np.float32)
    assert np.array_equal(back.view(np.uint32), ref.view(np.uint32))
    q = str(tmp_path / "e.dat")
    open(q, "w").write(" 1.25000000e+00   -3.5e-3 7\n4e2 5.000000 -0.000001")      # no trailing newline on the last row
    assert np.array_equal(topkrec.dat_read(q), np.array([[1.25, -3.5e-3, 7], [400, 5, -1e-6]], np.float32))
    open(q, "w").write("1 2 3\n4 5\n")
    with pytest.raises(topkrec.TkrError):
        topkrec.dat_read(q)
    topkrec.dat_write(q, np.zeros((0, 4), np.float32))
    assert open(q).read() == "" and topkrec.dat_read(q).shape[0] == 0
    # float64 input (CER's E, single/cer.py:81-85) is formatted from the doubles, exactly like the reference's "'%f ' % x"
    x64 = rng.standard_normal((300, 23)) * rng.choice([1e-6, 1.0, 1e5], (300, 23))
    topkrec.dat_write(q, x64)
    assert open(q).read() == "".join("".join("%f " % v for v in row) + "\n" for row in x64)
    assert open(q).read() != "".join("".join("%f " % v for v in row) + "\n" for row in x64.astype(np.float32).astype(np.float64))
    with pytest.raises(topkrec.TkrError):
        topkrec.dat_read(str(tmp_path / "missing.dat"))


def test_native_rating_paths_match_python_routes(mini, tmp_path):
    """rated_csr_from_files / test_lines_from_files (tkr_ratings_parse + numpy) == the dict-based routes that mirror
    the reference; incl. a user listed twice (the reference keeps the LAST line), unknown ids and a bare-uid line"""
    from oracle import evaluate_ref
    uid_file = os.path.join(mini, "uid")
    uids = utils.get_id_dict_from_file(uid_file)
    for sc in ("im", "om", "all"):
        idl = os.path.join(mini, "f0te.%s.idl" % sc)
        teids = utils.get_id_dict_from_file(idl)
        browsed, _ = utils.get_history_from_file(os.path.join(mini, "f0tr.txt"))
        p0, i0 = utils.rated_csr(uids, browsed, teids)
        p1, i1 = utils.rated_csr_from_files(uid_file, os.path.join(mini, "f0tr.txt"), idl, len(uids))
        assert np.array_equal(p0, p1) and np.array_equal(i0, i1)
        rows, ptr, idx = utils.test_lines_from_files(uid_file, os.path.join(mini, "f0te.%s.txt" % sc), idl)
        want_rows, want = [], []
        for line in open(os.path.join(mini, "f0te.%s.txt" % sc)):
            t = line.strip().split(",")
            likes = sorted({teids[x.split(":")[0]] for x in t[1:] if int(x.split(":")[1]) == 1})
            if likes:
                want_rows.append(uids[t[0]]); want.append(likes)
        assert rows.tolist() == want_rows and [idx[ptr[l]:ptr[l + 1]].tolist() for l in range(len(want))] == want
    # synthetic corner cases
    d = tmp_path
    (d / "uid").write_text("a\nb\nc\n"); (d / "vid").write_text("x\ny\nz\nw\n")
    (d / "tr.txt").write_text("a,x:1,y:0\nzz,x:1\nb\nc,w:1,q:1,z:0\na,z:1\n")
    browsed, _ = utils.get_history_from_file(str(d / "tr.txt"))
    u = utils.get_id_dict_from_file(str(d / "uid")); v = utils.get_id_dict_from_file(str(d / "vid"))
    p0, i0 = utils.rated_csr(u, browsed, v)
    p1, i1 = utils.rated_csr_from_files(str(d / "uid"), str(d / "tr.txt"), str(d / "vid"), 3)
    assert np.array_equal(p0, p1) and np.array_equal(i0, i1) and i1.tolist() == [2, 2, 3]
    n, tr_users, tr_data = utils.positives_from_files(str(d / "uid"), str(d / "vid"), str(d / "tr.txt"))
    data = utils.get_data_from_file(str(d / "tr.txt"), u, v)
    assert n == len(data) == 3 and tr_users == [0, 2] and tr_data == {0: [0, 2], 2: [3]}


def test_host_sampler_replays_reference_stream(golden, mini):
    z = np.load(os.path.join(golden, "sampler.npz"))
    m = single.BPR(k=8)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    np.random.seed(int(z["seed"]))
    gen = m._uniform_user_sampling(int(z["batch"]))
    for t in range(z["ub"].shape[0]):
        ub, ib, jb = next(gen)
        assert np.array_equal(ub, z["ub"][t]) and np.array_equal(ib, z["ib"][t]) and np.array_equal(jb, z["jb"][t])


def test_history_and_csr(mini):
    browsed, counter = utils.get_history_from_file(os.path.join(mini, "f0tr.txt"))
    uids = utils.get_id_dict_from_file(os.path.join(mini, "uid"))
    teids = utils.get_id_dict_from_file(os.path.join(mini, "f0te.im.idl"))
    indptr, idx = utils.rated_csr(uids, browsed, teids)
    assert indptr.shape == (301,) and indptr[-1] == idx.shape[0]
    from oracle import evaluate_ref
    rp, ri = evaluate_ref.rated_csr(uids, evaluate_ref.load_history(os.path.join(mini, "f0tr.txt")), teids)
    assert np.array_equal(indptr, rp) and np.array_equal(idx, ri)
    for r in range(300):
        seg = idx[indptr[r]:indptr[r + 1]]
        assert (np.diff(seg) > 0).all()
    assert sum(counter.values()) > 0


def test_positives_csr_sorted(mini):
    m = single.BPR(k=8)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    indptr, idx = utils.positives_csr(m.tr_users, m.tr_data, m.n_users)
    for u in m.tr_users:
        assert sorted(m.tr_data[u]) == idx[indptr[u]:indptr[u + 1]].tolist()
    assert indptr[-1] == m.epoch_sample_limit


def test_class_surface_matches_reference():
    import inspect
    sig = inspect.signature(single.BPR.__init__)
    names = list(sig.parameters)[1:8]
    assert names == ["k", "lambda_u", "lambda_i", "lambda_j", "lambda_b", "lr", "mode"]
    d = {n: sig.parameters[n].default for n in names[1:]}
    assert d == {"lambda_u": 2.5e-3, "lambda_i": 2.5e-3, "lambda_j": 2.5e-4, "lambda_b": 0, "lr": 1.0e-4, "mode": "l2"}
    t = inspect.signature(single.BPR.train)
    assert list(t.parameters)[1:] == ["sampling", "epochs", "batch_size", "epoch_sample_limit", "model_path"]
    assert [t.parameters[n].default for n in list(t.parameters)[1:]] == ["user uniform", 5, 256, None, None]
    v = inspect.signature(single.VBPR.__init__)
    assert list(v.parameters)[1:10] == ["k", "d", "lambda_u", "lambda_i", "lambda_j", "lambda_b", "lambda_e", "lr", "mode"]
    for meth in ("load_training_data", "load_content_data", "build_graph", "train", "export_model", "export_embeddings",
                 "import_model", "import_embeddings"):
        assert hasattr(single.REC, meth)
    m = single.BPR(k=50)
    for attr in ("fue", "fie", "fib", "uids", "iids", "n_users", "n_items", "tr_data", "tr_users", "epoch_sample_limit",
                 "k", "lu", "li", "lj", "lb", "lr", "mode"):
        assert hasattr(m, attr)


def test_export_import_embeddings_roundtrip(tmp_path, mini):
    m = single.BPR(k=4)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    rng = np.random.default_rng(0)
    m.fue = rng.standard_normal((m.n_users, 4)).astype(np.float32)
    m.fie = rng.standard_normal((m.n_items, 4)).astype(np.float32)
    m.fib = rng.standard_normal((m.n_items, 1)).astype(np.float32)
    out = str(tmp_path / "embed" / "bpr")          # parent missing: we makedirs (superset of D-6)
    m.export_embeddings(out)
    assert sorted(os.listdir(out)) == ["final-B.dat", "final-U.dat", "final-V.dat"]
    m2 = single.BPR(k=4); m2.uids, m2.iids = m.uids, m.iids
    m2.import_embeddings(out)
    for a in ("fue", "fie", "fib"):
        assert np.abs(getattr(m2, a) - getattr(m, a)).max() <= 1e-6
        assert getattr(m2, a).shape == getattr(m, a).shape


def test_load_content_data(tmp_path, mini):
    import pickle
    import scipy.sparse as ss
    m = single.VBPR(k=8, d=6)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    F = ss.random(160, 6, density=0.3, format="lil", dtype=np.float32, random_state=1)
    p = tmp_path / "meta.pkl"
    pickle.dump(F, open(p, "wb"))
    sub = tmp_path / "ids"                           # content rows follow a different id order
    order = list(reversed(list(m.iids)))[:100]
    sub.write_text("".join(i + "\n" for i in order))
    m.load_content_data(str(p), str(sub))
    dense = F.toarray()
    for pos, iid in enumerate(order):
        assert np.array_equal(m.feat[m.iids[iid]], dense[pos])
    missing = [i for i in m.iids if i not in set(order)]
    assert all((m.feat[m.iids[i]] == 0).all() for i in missing)


def test_als_plan_tiles_every_row():
    """topkrec.als_build_plan: segments of a row are consecutive, cover its positives exactly once, split rows own
    consecutive partial slots, longest rows come first."""
    import topkrec
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 50, 200); cnt[5] = 1000; cnt[9] = 64; cnt[11] = 0
    indptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    for seg in (16, 64, 4096):
        plan, n_slots = topkrec.als_build_plan(indptr, seg)
        seen = np.zeros(int(indptr[-1]), np.int32)
        for r, off, ln in zip(plan["seg_row"], plan["seg_off"], plan["seg_len"]):
            assert indptr[r] <= off and off + ln <= indptr[r + 1] and 0 <= ln <= seg
            seen[off:off + ln] += 1
        assert np.all(seen == 1)
        assert sorted(set(plan["seg_row"].tolist())) == list(range(200))              # empty rows keep one (empty) segment
        first_len = cnt[plan["seg_row"]]
        assert np.all(np.diff(first_len) <= 0)                                        # longest first
        multi = set(np.flatnonzero(cnt > seg).tolist())
        assert set(plan["multi_row"].tolist()) == multi and n_slots == int((plan["seg_slot"] >= 0).sum())
        for r, s0, ns, tot in zip(plan["multi_row"], plan["multi_slot0"], plan["multi_nslots"], plan["multi_total"]):
            slots = plan["seg_slot"][plan["seg_row"] == r]
            assert slots.tolist() == list(range(s0, s0 + ns)) and tot == cnt[r] and ns == -(-cnt[r] // seg)
        single = plan["seg_slot"][~np.isin(plan["seg_row"], plan["multi_row"])]
        assert np.all(single == -1)
    with pytest.raises(ValueError):
        topkrec.als_build_plan(np.array([0, 3, 2]), 16)


def test_als_abi_argument_errors_without_gpu():
    import ctypes as C
    import topkrec
    from topkrec._lib import tkr_als_cfg, tkr_als_plan
    L = topkrec.lib()
    assert L.tkr_als_partial_bytes(300, 1) == 0 and L.tkr_als_gram_workspace_bytes(0) == 0
    assert L.tkr_als_partial_bytes(256, 2) >= 2 * (10 * 4096 + 256) * 4
    assert L.tkr_als_gram(None, 8, None, 0, 1.0, 0.0, None, None, 0, None) == -1 and b"null pointer" in L.tkr_last_error()
    cfg, plan = tkr_als_cfg(512, 1, 0.01, 0, 0.01, 0, 0), tkr_als_plan()
    assert L.tkr_als_solve_rows(C.byref(cfg), C.byref(plan), 1, 1, None, 1, None, None, None, 0, None) == -1
    assert b"outside [1,256]" in L.tkr_last_error()


def test_wmf_cer_class_surface():
    """constructor defaults and attributes of single/wmf.py:11-31, single/cer.py:17-22"""
    import inspect
    from single import WMF, CER, REC
    w, c = WMF(8), CER(8, 5)
    assert (w.k, w.lu, w.lv, w.a, w.b) == (8, 0.01, 0.01, 1, 0.01)
    assert (c.k, c.d, c.lu, c.lv, c.le, c.a, c.b) == (8, 5, 0.01, 10, 10e3, 1, 0.01) and c.E is None
    assert issubclass(CER, WMF) and issubclass(WMF, REC)
    for cls in (WMF, CER):
        sig = inspect.signature(cls.train)
        assert list(sig.parameters)[1:] == ["max_iter", "tol", "model_path"]
        assert sig.parameters["max_iter"].default == 200 and sig.parameters["tol"].default == 1e-4
    for attr in ("uids", "n_users", "usm", "iids", "n_items", "ism", "n_ratings", "u_rated", "i_rated", "fue", "fie"):
        assert hasattr(w, attr)


def _mini_without_unknown_user(mini, tmp_path):
    tr = tmp_path / "tr.txt"
    tr.write_text("".join(ln for ln in open(os.path.join(mini, "f0tr.txt")) if not ln.startswith("99999,")))
    return str(tr)


def test_wmf_loader_matches_reference(golden, mini, tmp_path):
    """WMF.load_training_data (host only) against the reference's own loader (wmf.py:33-56), whose usm / ism lists and
    seeded uniform(0,1) start are stored in tests/golden/als_cer.npz by make_golden_als.py."""
    from single import CER
    g = np.load(os.path.join(golden, "als_cer.npz"))
    m = CER(k=int(g["k"]), d=int(g["d_feat"]))
    with pytest.raises(KeyError):                                           # wmf.py:51: unknown uid inside a positive pair
        m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    np.random.seed(77)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), _mini_without_unknown_user(mini, tmp_path))
    u_ptr, u_idx, i_ptr, i_idx = m._csr
    assert np.array_equal(np.diff(u_ptr), g["u_cnt"]) and np.array_equal(u_idx, g["u_idx"])      # usm, file order kept
    assert np.array_equal(np.diff(i_ptr), g["i_cnt"]) and np.array_equal(i_idx, g["i_idx"])      # ism
    assert m.usm[5] == g["u_idx"][u_ptr[5]:u_ptr[6]].tolist() and m.ism[7] == g["i_idx"][i_ptr[7]:i_ptr[8]].tolist()
    assert m.u_rated == np.flatnonzero(g["u_cnt"] > 0).tolist() and m.i_rated == np.flatnonzero(g["i_cnt"] > 0).tolist()
    assert m.n_ratings == m.n_users * m.n_items
    assert np.array_equal(m.fue, g["fue0"]) and np.array_equal(m.fie, g["fie0"])                 # same RNG draws, same order


def test_cer_model_files_roundtrip(tmp_path, mini):
    """export_embeddings writes final-U/V.dat + final-E.dat (cer.py:81-85), import_embeddings reads them back (cer.py:75-79)."""
    from single import CER
    m = CER(k=6, d=4)
    np.random.seed(1)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), _mini_without_unknown_user(mini, tmp_path))
    m.E = np.random.randn(4, 6)
    path = str(tmp_path / "cer")
    m.export_embeddings(path)
    assert sorted(os.listdir(path)) == ["final-E.dat", "final-U.dat", "final-V.dat"]
    r = CER(k=6, d=4)
    r.uids, r.iids = m.uids, m.iids
    r.import_embeddings(path)
    assert np.allclose(r.fue, m.fue, atol=5e-7) and np.allclose(r.fie, m.fie, atol=5e-7) and np.allclose(r.E, m.E, atol=5e-7)
    assert r.E.shape == (4, 6)


def test_single_exports_the_reference_names():
    """single/__init__.py:1-9 exports REC, BPR, VBPR, WMF, DPM, CER, ENCODER, MLP; DPM/MLP fail where the reference's do."""
    import single
    for name in ("REC", "BPR", "VBPR", "WMF", "DPM", "CER", "ENCODER", "MLP"):
        assert hasattr(single, name) and name in single.__all__
    with pytest.raises(TypeError):
        single.MLP(8, 4)
    m = single.DPM(k=8, d=4)
    assert (m.k, m.d, m.lv, m.le) == (8, 4, 10, 10e3)
    with pytest.raises(TypeError):
        m.train(single.MLP, max_iter=1)

"""GPU parity tests of the VBPR step (tkr_vbpr_step through the C ABI) against the numpy oracle
(oracle/bpr_ref.py::vbpr_step).  Tolerance 1e-4 relative, as for BPR."""
import os

import numpy as np
import pytest
import torch

import topkrec
from oracle import bpr_ref

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _dev_state(st, F, k):
    """oracle state (UR, UC, IR, rb, E, c) -> engine layout (U = [ur|uc], V = [ir | F.E], bsum = rb + F.c)."""
    h = k // 2
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    ni = st["IR"].shape[0]
    d = {"U": t(np.concatenate([st["UR"], st["UC"]], 1)), "V": t(np.concatenate([st["IR"], np.zeros((ni, h), np.float32)], 1)),
         "rb": t(st["rb"]), "bsum": t(np.zeros(ni)), "E": t(st["E"]), "c": t(st["c"])}
    d["msU"] = t(np.concatenate([st["msUR"], st["msUC"]], 1)); d["msV"] = t(np.concatenate([st["msIR"], np.ones((ni, h), np.float32)], 1))
    d["msrb"] = t(st["msrb"]); d["msE"] = t(st["msE"]); d["msc"] = t(st["msc"])
    return d


def _run(nu, ni, k, dF, B, steps, seed, sparse_feat=False, **cfg_kw):
    rng = np.random.default_rng(seed)
    st = bpr_ref.new_vbpr_state(nu, ni, k, dF, rng)
    st["rb"] = (0.01 * rng.standard_normal(ni)).astype(np.float32)
    st["c"] = (0.001 * rng.standard_normal(dF)).astype(np.float32)
    st["E"] = (st["E"] * (1 + 0.5 * rng.standard_normal(st["E"].shape))).astype(np.float32)
    if sparse_feat:   # binary bag-of-words rows like the shipped meta.pkl, some empty
        F = (rng.random((ni, dF)) < 0.02).astype(np.float32); F[::7] = 0
    else:             # |N(0,1)| row-normalised
        F = np.abs(rng.standard_normal((ni, dF))).astype(np.float32); F /= np.linalg.norm(F, axis=1, keepdims=True)
    u = rng.integers(0, nu, B * steps).astype(np.int32)
    p = 1.0 / np.arange(1, ni + 1); p /= p.sum()
    i = rng.choice(ni, B * steps, p=p).astype(np.int32); j = rng.integers(0, ni, B * steps).astype(np.int32)
    ocfg = bpr_ref.BprCfg(**cfg_kw)
    cfg = topkrec.VbprCfg(nu, ni, k, dF, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lambda_e, ocfg.lr, ocfg.mode, ocfg.optimizer)
    d = _dev_state(st, F, k)
    Fd = torch.from_numpy(F).cuda()
    ws = topkrec.vbpr_workspace(cfg, B)
    loss = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.vbpr_project(cfg, d, Fd)
    topkrec.vbpr_step(cfg, d, Fd, torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, loss)
    ref_loss = np.array([bpr_ref.vbpr_step(st, F, u[t * B:(t + 1) * B], i[t * B:(t + 1) * B], j[t * B:(t + 1) * B], ocfg) for t in range(steps)])
    h = k // 2
    got = {n: v.cpu().numpy() for n, v in d.items()}
    pairs = {"UR": got["U"][:, :h], "UC": got["U"][:, h:], "IR": got["V"][:, :h], "rb": got["rb"], "E": got["E"], "c": got["c"],
             "msUR": got["msU"][:, :h], "msUC": got["msU"][:, h:], "msIR": got["msV"][:, :h], "msrb": got["msrb"], "msE": got["msE"], "msc": got["msc"]}
    for n, g in pairs.items():
        assert _rel(g, st[n]) <= REL_TOL, (n, _rel(g, st[n]))
    assert np.allclose(loss.cpu().numpy(), ref_loss, rtol=1e-4)
    # export identity (vbpr.py:124-126): the engine state IS (fue, fie, fib)
    fue, fie, fib = bpr_ref.vbpr_export(st, F)
    assert _rel(got["U"], fue) <= REL_TOL and _rel(got["V"], fie) <= REL_TOL and _rel(got["bsum"], fib.ravel()) <= REL_TOL
    assert int(ws.count_nonzero().item()) <= 4 * (min(B, nu) + min(2 * B, ni)), "accumulators must be re-zeroed (only the row lists may be stale)"


@pytest.mark.parametrize("k,dF", [(128, 512), (50, 300), (16, 70), (64, 1030)])
def test_vbpr_step_matches_oracle(k, dF):
    _run(300, 200, k, dF, 256, 8, seed=k + dF)


def test_vbpr_sparse_binary_features_like_meta_pkl():
    _run(400, 300, 50, 2000, 256, 10, seed=1, sparse_feat=True)


def test_vbpr_large_batch_dense_mode():
    _run(500, 250, 128, 256, 4096, 4, seed=2)


@pytest.mark.parametrize("kw", [dict(lambda_e=0.01, lambda_b=0.05), dict(mode="l1", lambda_e=0.001, lambda_b=0.01), dict(optimizer="sgd", lambda_e=0.01)])
def test_vbpr_modes(kw):
    _run(200, 150, 32, 128, 256, 10, seed=3, **kw)


def test_vbpr_class_end_to_end(tmp_path, mini):
    """train.py:11-16 on the mini fixture: load -> content -> train -> export -> warm-start train; the export
    identity fue.fie^T + fib == x_ui lets evaluate score VBPR models without features."""
    import pickle
    import scipy.sparse as ss
    import single
    np.random.seed(5)
    m = single.VBPR(k=16, d=120, seed=7)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    F = ss.random(160, 120, density=0.1, format="lil", dtype=np.float32, random_state=2)
    pickle.dump(F, open(tmp_path / "meta.pkl", "wb"))
    m.load_content_data(str(tmp_path / "meta.pkl"), os.path.join(mini, "vid"))
    m.train(epochs=2, batch_size=64, epoch_sample_limit=10e2)
    assert m.fue.shape == (300, 16) and m.fie.shape == (160, 16) and m.fib.shape == (160, 1)
    assert len(m.losses) == 2 * (1000 // 64) and np.isfinite(m.losses).all()
    st = {n: v.cpu().numpy() for n, v in m._state.items()}
    assert np.allclose(m.fie[:, 8:], m.feat @ st["E"], rtol=1e-4, atol=1e-7)
    assert np.allclose(m.fib.ravel(), st["rb"] + m.feat @ st["c"], rtol=1e-4, atol=1e-7)
    m.export_embeddings(str(tmp_path / "vbpr"))
    m.train(epochs=1, batch_size=64, epoch_sample_limit=10e2, model_path=str(tmp_path / "vbpr"))
    assert np.isfinite(m.fue).all()

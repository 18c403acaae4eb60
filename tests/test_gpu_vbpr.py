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


def _set_tc(mode):
    import ctypes
    L = topkrec.lib()
    L.tkr_debug_set_vbpr_tc_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_vbpr_tc_mode.restype = None
    L.tkr_debug_set_vbpr_tc_mode(mode)


def _core_ws_bytes(cfg, B):
    """size of the step workspace without the tensor-core route's buffers (F^T and the pre-split B operands stay non-zero)"""
    _set_tc(0)
    n = topkrec.lib().tkr_vbpr_workspace_bytes(cfg.ptr, B)
    _set_tc(-1)
    return n


def _run(nu, ni, k, dF, B, steps, seed, sparse_feat=False, hot=False, pairwise=False, **cfg_kw):
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
    cfg = topkrec.VbprCfg(nu, ni, k, dF, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lambda_e, ocfg.lr, ocfg.mode, ocfg.optimizer, pairwise=pairwise)
    d = _dev_state(st, F, k)
    Fd = torch.from_numpy(F).cuda()
    ws = topkrec.vbpr_workspace(cfg, B)
    if hot:
        topkrec.vbpr_set_hot_items(cfg, B, ws, topkrec.popular_items(i, ni))
    loss = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.vbpr_project(cfg, d, Fd)
    topkrec.vbpr_step(cfg, d, Fd, torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, loss)
    ref_loss = np.array([bpr_ref.vbpr_step(st, F, u[t * B:(t + 1) * B], i[t * B:(t + 1) * B], j[t * B:(t + 1) * B], ocfg, pairwise=pairwise) for t in range(steps)])
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
    plain = topkrec.VbprCfg(nu, ni, k, dF, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lambda_e, ocfg.lr, ocfg.mode, ocfg.optimizer)
    core = ws[:_core_ws_bytes(plain, B)]          # (without the tensor-core and pairwise scratch regions, which stay non-zero)
    # (stale by design: the touched-row lists of the step kernels, and for small batches the projected-row list + the triples drawn ahead)
    assert int(core.count_nonzero().item()) <= 4 * (min(B, nu) + min(2 * B, ni)) + 20 * B + 16, "accumulators must be re-zeroed (only row lists / staged triples may be stale)"
    return d


@pytest.mark.parametrize("k,dF", [(128, 512), (50, 300), (16, 70), (64, 1030)])
def test_vbpr_step_matches_oracle(k, dF):
    _run(300, 200, k, dF, 256, 8, seed=k + dF)


def test_vbpr_sparse_binary_features_like_meta_pkl():
    _run(400, 300, 50, 2000, 256, 10, seed=1, sparse_feat=True)


def test_vbpr_large_batch_dense_mode():
    _run(500, 250, 128, 256, 4096, 4, seed=2)


@pytest.mark.parametrize("shape", [(500, 250, 128, 256, 4096, 4), (300, 200, 50, 300, 256, 8)])
def test_vbpr_hot_items_privatised(shape):
    """popular item rows (rating part, content gradient W, wq) summed per thread block in shared memory: same sums"""
    nu, ni, k, dF, B, steps = shape
    _run(nu, ni, k, dF, B, steps, seed=21, hot=True, lambda_b=0.01)


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


def test_vbpr_tensor_core_gemms_fp32_level_accuracy():
    """the two content GEMMs on tcgen05 kind::tf32 with the 3-term split, at the C3 shape (10 000 x 4096, h = 64): the
    projection [F.E | F.c] and one step's dE / dc against fp64 -- the error must be fp32-level (plain TF32 would be ~1e-3)"""
    rng = np.random.default_rng(7)
    nu, ni, k, dF, B = 2000, 10000, 128, 4096, 1 << 15
    h = k // 2
    F = np.abs(rng.standard_normal((ni, dF), dtype=np.float32)); F /= np.linalg.norm(F, axis=1, keepdims=True)
    st = bpr_ref.new_vbpr_state(nu, ni, k, dF, rng)
    st["E"] = (0.05 * rng.standard_normal((dF, h))).astype(np.float32)
    st["c"] = (0.05 * rng.standard_normal(dF)).astype(np.float32)
    st["rb"] = (0.01 * rng.standard_normal(ni)).astype(np.float32)
    cfg = topkrec.VbprCfg(nu, ni, k, dF)
    d = _dev_state(st, F, k)
    Fd = torch.from_numpy(F).cuda()
    ws = topkrec.vbpr_workspace(cfg, B)
    assert ws.numel() > _core_ws_bytes(cfg, B), "this shape must take the tensor-core route"
    z = torch.zeros(B, dtype=torch.int32, device="cuda")
    topkrec.vbpr_step(cfg, d, Fd, z, z, z, B, 0, ws, None)                                   # zero steps: just the projection
    P = F.astype(np.float64) @ st["E"].astype(np.float64)
    q = st["rb"].astype(np.float64) + F.astype(np.float64) @ st["c"].astype(np.float64)
    got = d["V"].cpu().numpy()[:, h:]
    err = np.abs(got - P).max() / np.abs(P).max()
    errq = np.abs(d["bsum"].cpu().numpy() - q).max() / np.abs(q).max()
    # measured 8.6e-6 / 7.0e-6 (K = 4096 same-sign products: the tensor core's own accumulation truncates; rotating over
    # four accumulators keeps that chain short).  Plain TF32 operands would sit at ~1e-3.
    assert err <= 2e-5 and errq <= 2e-5, (err, errq)


@pytest.mark.parametrize("shape", [(500, 250, 128, 256, 4096, 4), (3000, 1000, 50, 1000, 1 << 14, 3), (800, 129, 16, 68, 2048, 5)])
def test_vbpr_tensor_core_route_equals_cuda_core_route(shape):
    """the same stream through both routes of the content GEMMs (ragged M / K tiles, h + 1 not a multiple of 16,
    split-K gradient): states agree far inside the parity bar, and both meet it against the oracle"""
    nu, ni, k, dF, B, steps = shape
    outs = []
    for mode in (0, -1):
        _set_tc(mode)
        try:
            outs.append({n: v.cpu().numpy() for n, v in _run(nu, ni, k, dF, B, steps, seed=9).items()})
        finally:
            _set_tc(-1)
    for n in outs[0]:
        assert _rel(outs[1][n], outs[0][n]) <= 2e-5, n


def test_vbpr_data_parallel_halves_equal_one_big_batch():
    """tkr_vbpr_grad on two user-partitioned half batches + summed [GV|Gb|tchV] and [GE|Gc] regions + tkr_vbpr_apply on both
    == one step over the union batch (what 2 ranks + two all-reduces compute), for 3 steps."""
    rng = np.random.default_rng(31)
    nu, ni, k, dF, B, steps = 600, 300, 64, 256, 2048, 3
    h = k // 2
    st = bpr_ref.new_vbpr_state(nu, ni, k, dF, rng)
    st["rb"] = (0.01 * rng.standard_normal(ni)).astype(np.float32)
    st["c"] = (0.001 * rng.standard_normal(dF)).astype(np.float32)
    F = np.abs(rng.standard_normal((ni, dF))).astype(np.float32); F /= np.linalg.norm(F, axis=1, keepdims=True)
    Fd = torch.from_numpy(F).cuda()
    ocfg = bpr_ref.BprCfg(lambda_e=0.01, lambda_b=0.01)
    cfg = topkrec.VbprCfg(nu, ni, k, dF, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lambda_e, ocfg.lr, ocfg.mode, ocfg.optimizer)
    ranks = []
    for r in range(2):
        d = _dev_state(st, F, k)
        ws = topkrec.vbpr_workspace(cfg, B)
        topkrec.vbpr_project(cfg, d, Fd)
        ranks.append((d, ws, topkrec.vbpr_grad_views(cfg, B, ws)))
    for t in range(steps):
        u = rng.integers(0, nu, 2 * B).astype(np.int32); i = rng.integers(0, ni, 2 * B).astype(np.int32); j = rng.integers(0, ni, 2 * B).astype(np.int32)
        u[:B] = u[:B] // 2 * 2; u[B:] = u[B:] // 2 * 2 + 1              # rank r owns users u % 2 == r
        loss = [torch.zeros(1, device="cuda") for _ in range(2)]
        for r, (d, ws, _) in enumerate(ranks):
            sl = slice(r * B, (r + 1) * B)
            topkrec.vbpr_grad(cfg, d, Fd, torch.from_numpy(u[sl]).cuda(), torch.from_numpy(i[sl]).cuda(), torch.from_numpy(j[sl]).cuda(), B, ws, loss[r],
                              data_parallel=True)
        for v in range(2):                                                # the two all-reduces
            total = ranks[0][2][v] + ranks[1][2][v]
            ranks[0][2][v].copy_(total); ranks[1][2][v].copy_(total)
        for d, ws, _ in ranks:
            topkrec.vbpr_apply(cfg, d, B, ws, None, data_parallel=True)
        ref_loss = bpr_ref.vbpr_step(st, F, u, i, j, ocfg)
        data_loss = (loss[0] + loss[1]).item()                            # (the halves add the batch terms; the tiny E / c regularisers come with apply)
        assert abs(data_loss - ref_loss) / ref_loss < 1e-2
    a, b = ranks[0][0], ranks[1][0]
    for n in ("V", "rb", "E", "c", "msV", "msrb", "msE", "msc"):          # replicas identical ...
        assert torch.equal(a[n][:, :h] if n in ("V", "msV") else a[n], b[n][:, :h] if n in ("V", "msV") else b[n]), n
    got = {n: v.cpu().numpy() for n, v in a.items()}
    for n, g in (("IR", got["V"][:, :h]), ("rb", got["rb"]), ("E", got["E"]), ("c", got["c"]), ("msIR", got["msV"][:, :h]), ("msE", got["msE"]), ("msc", got["msc"])):
        assert _rel(g, st[n]) <= REL_TOL, (n, _rel(g, st[n]))             # ... and equal to the union batch
    U = got["U"].copy(); U[1::2] = b["U"].cpu().numpy()[1::2]             # each rank updated only its own users
    assert _rel(U[:, :h], st["UR"]) <= REL_TOL and _rel(U[:, h:], st["UC"]) <= REL_TOL


@pytest.mark.parametrize("shape,kw", [((300, 200, 64, 300, 256, 6), {}), ((200, 150, 32, 128, 64, 10), dict(lambda_e=0.01, lambda_b=0.05)),
                                      ((400, 300, 50, 500, 700, 3), dict(mode="l1", lambda_b=0.01)), ((300, 200, 16, 70, 4096, 2), {})])
def test_vbpr_graph_as_written_matches_oracle(shape, kw):
    """tkr_vbpr_cfg.pairwise = 1: the reference's graph exactly as written -- vbpr.py:61 broadcasts x to [B, B] (D-14), every
    embedding-side weight is a column sum and every bias-side weight a row sum of sigma(-x) -- against the oracle whose pairwise
    form agrees with the literal torch transcription of the graph (tests/test_oracle.py)"""
    nu, ni, k, dF, B, steps = shape
    _run(nu, ni, k, dF, B, steps, seed=51, pairwise=True, **kw)


def test_vbpr_graph_as_written_fused_sampler_and_limits(mini):
    """the pairwise pre-pass needs the batch's triples before the gradient kernel: with the fused sampler they are drawn ahead
    (same draws); batches beyond 4096 are refused"""
    from test_gpu_bpr import _mini_tables
    tr_users, tr_data, indptr, idx, nu, ni = _mini_tables(mini)
    rng = np.random.default_rng(52)
    k, dF, B, steps = 16, 40, 128, 5
    st = bpr_ref.new_vbpr_state(nu, ni, k, dF, rng)
    F = np.abs(rng.standard_normal((ni, dF))).astype(np.float32)
    cfg = topkrec.VbprCfg(nu, ni, k, dF, pairwise=True)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=5)
    Fd = torch.from_numpy(F).cuda()
    a, b = _dev_state(st, F, k), _dev_state(st, F, k)
    ws = topkrec.vbpr_workspace(cfg, B)
    la, lb = torch.empty(steps, device="cuda"), torch.empty(steps, device="cuda")
    topkrec.vbpr_project(cfg, a, Fd); topkrec.vbpr_project(cfg, b, Fd)
    topkrec.vbpr_step(cfg, a, Fd, None, None, None, B, steps, ws, la, sampler=smp, first_draw=77)
    u, i, j = topkrec.bpr_sample(smp, 77, B * steps)
    topkrec.vbpr_step(cfg, b, Fd, u, i, j, B, steps, ws, lb)
    for n in a:
        assert _rel(a[n].cpu().numpy(), b[n].cpu().numpy()) <= 2e-6, n
    assert np.allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=1e-5)
    with pytest.raises(topkrec.TkrError, match="O\\(B\\^2\\)"):
        topkrec.vbpr_workspace(cfg, 8192)


def test_vbpr_small_batch_projects_only_touched_rows():
    """B = 256 on a table of 3 000 items: only the <= 512 item rows of a batch are re-projected per step (the export is refreshed
    for every row at the end); explicit triples and the fused sampler (triples drawn ahead) give the same result as the oracle"""
    d = _run(700, 3000, 64, 300, 256, 12, seed=71, lambda_e=0.01, lambda_b=0.01)
    assert d is not None

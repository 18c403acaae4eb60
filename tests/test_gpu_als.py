"""GPU parity of the ALS path (SURVEY.md 8(f) NEXT-1; reference single/cer.py:24-73, single/wmf.py:61-101) through
the C ABI (tkr_als_gram / tkr_als_solve_rows) against oracle/als_ref.py and the reference-generated golden
tests/golden/als_cer.npz.

Tolerance (stated, north star: "learned U/V within 1e-4 relative"): max |X - X_ref| / max |X_ref| <= 1e-4 per factor.
The reference forms each normal matrix in fp32 and solves it in fp64 LAPACK; the kernel forms it in fp32 (another
summation order) and factors it in fp32 (L D L^T), so agreement is bounded by cond(A) * 2^-24, not by bit-equality.
"""
import os

import numpy as np
import pytest
import torch

import topkrec
from oracle import als_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(x, ref):
    return float(np.abs(np.asarray(x, np.float64) - ref).max() / np.abs(ref).max())


def exact64_step(X, Y, ptr, idx, rated, a, b, base_ridge, ridge, prior=None, solve_empty=False):
    """The same half-step with every matrix formed and solved in fp64 from the fp32 inputs: the yardstick for the
    reference's OWN rounding (it forms each matrix in fp32).  With the reference's uniform(0,1) start (wmf.py:55-56)
    the systems reach cond ~ 3e4 and the reference sits ~1.7e-4 from this solution -- two fp32 summation orders of
    the same Gram matrix already differ by more than 1e-4 there -- so on such data the bound is
    max(1e-4, 2 x the reference's distance from exact); on well-conditioned data it is the plain 1e-4."""
    Y64 = Y.astype(np.float64)
    out = X.astype(np.float64).copy()
    k = Y.shape[1]
    XX = b * Y64[rated].T @ Y64[rated] + base_ridge * np.eye(k)
    for r in range(X.shape[0]):
        pos = idx[ptr[r]:ptr[r + 1]]
        if len(pos) == 0 and not solve_empty:
            continue
        Yi = Y64[pos]
        rhs = a * Yi.sum(0) + (ridge * prior[r].astype(np.float64) if prior is not None else 0.0)
        out[r] = np.linalg.solve(XX + (a - b) * Yi.T @ Yi + ridge * np.eye(k), rhs)
    return out


def bound(ref, exact):
    return max(TOL, 2.0 * rel(ref, exact))


def random_csr(rng, n_rows, n_cols, max_len, empty_frac=0.1, long_rows=()):
    cnt = rng.integers(1, max_len + 1, n_rows)
    cnt[rng.random(n_rows) < empty_frac] = 0
    for r, n in long_rows:
        cnt[r] = n
    indptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(cnt, out=indptr[1:])
    idx = rng.integers(0, n_cols, int(indptr[-1])).astype(np.int32)          # duplicates allowed (wmf.py:51 appends blindly)
    return indptr, idx


@pytest.mark.parametrize("d", [1, 7, 50, 64, 100, 128, 130, 192, 256])
def test_gram_matches_numpy(d):
    rng = np.random.default_rng(d)
    Y = rng.random((3000, d)).astype(np.float32)
    for n in (0, 1, 17, 2999):
        rows = np.sort(rng.choice(3000, n, replace=False)).astype(np.int32)
        got = topkrec.als_gram(torch.from_numpy(Y).cuda(), torch.from_numpy(rows).cuda(), 0.01, 0.5).cpu().numpy()
        Yr = Y[rows].astype(np.float64)
        ref = 0.01 * Yr.T @ Yr + 0.5 * np.eye(d)
        assert rel(got, ref) <= 2e-6
        assert np.array_equal(got, got.T)


@pytest.mark.parametrize("init", ["uniform", "normal"])
@pytest.mark.parametrize("d,seg", [(50, 4096), (50, 40), (64, 33), (12, 16), (128, 4096), (128, 100), (192, 64), (256, 4096), (256, 48)])
def test_user_and_item_steps_match_oracle(d, seg, init):
    rng = np.random.default_rng(100 + d + seg)
    n_users, n_items = 220, 150
    u_ptr, u_idx = random_csr(rng, n_users, n_items - 10, 60, long_rows=((3, 300), (77, 129)))   # the last 10 items stay unrated
    users = np.repeat(np.arange(n_users), np.diff(u_ptr))
    by_i = np.argsort(u_idx, kind="stable")
    i_ptr = np.zeros(n_items + 1, np.int64)
    np.cumsum(np.bincount(u_idx, minlength=n_items), out=i_ptr[1:])
    i_idx = users[by_i].astype(np.int32)
    if init == "uniform":                      # the reference's start (wmf.py:55-56): ill-conditioned systems
        fue = rng.random((n_users, d)).astype(np.float32)
        fie = rng.random((n_items, d)).astype(np.float32)
    else:                                      # centred factors, as after training: the plain 1e-4 bound while d < the
                                               # number of rows behind the shared Gram (d >= 100 of 140 items: near-singular)
        fue = (0.3 * rng.standard_normal((n_users, d))).astype(np.float32)
        fie = (0.3 * rng.standard_normal((n_items, d))).astype(np.float32)
    Fe = (0.3 * rng.standard_normal((n_items, d))).astype(np.float32)
    a, b, lu, lv = 1.0, 0.01, 0.01, 10.0
    u_rated = np.flatnonzero(np.diff(u_ptr) > 0); i_rated = np.flatnonzero(np.diff(i_ptr) > 0)

    us = topkrec.AlsSide(u_ptr, u_idx, seg)
    its = topkrec.AlsSide(i_ptr, i_idx, seg)
    assert (seg >= 4096) == (us.n_slots == 0)
    U, V = torch.from_numpy(fue).cuda(), torch.from_numpy(fie).cuda()

    # user half-step (cer.py:36-46)
    ref_u = fue.copy()
    loss_ref = als_ref.user_step(ref_u, fie, u_ptr, u_idx, i_rated, a, b, lu)
    XX = topkrec.als_gram(V, its.rated_dev, b, lu)
    lr = topkrec.als_solve_rows(us, V, U, XX, a, b, 0.0, lu)
    torch.cuda.synchronize()
    tol = bound(ref_u, exact64_step(fue, fie, u_ptr, u_idx, i_rated, a, b, lu, 0.0)) if init == "uniform" or d >= 100 else TOL
    assert rel(U.cpu().numpy(), ref_u.astype(np.float64)) <= tol
    assert abs(float(lr.sum()) - loss_ref) <= 1e-5 * (abs(loss_ref) + a * u_idx.size)      # terms of size a*nnz cancel in the item loss
    empty = np.flatnonzero(np.diff(u_ptr) == 0)
    assert np.array_equal(U.cpu().numpy()[empty], fue[empty])                # users without positives keep their row

    # item half-step, CER flavour (cer.py:47-63): prior + unrated items solved
    U0 = torch.from_numpy(ref_u).cuda()
    ref_v = fie.copy()
    loss_ref = als_ref.item_step(ref_u, ref_v, i_ptr, i_idx, u_rated, a, b, lv, Fe)
    XXv = topkrec.als_gram(U0, us.rated_dev, b, 0.0)
    Vc = V.clone()
    lr = topkrec.als_solve_rows(its, U0, Vc, XXv, a, b, lv, lv, prior=torch.from_numpy(Fe).cuda(), solve_empty=True, item_loss=True)
    tol = bound(ref_v, exact64_step(fie, ref_u, i_ptr, i_idx, u_rated, a, b, 0.0, lv, Fe, True)) if init == "uniform" or d >= 100 else TOL
    assert rel(Vc.cpu().numpy(), ref_v.astype(np.float64)) <= tol
    assert abs(float(lr.sum()) - loss_ref) <= 1e-5 * (abs(loss_ref) + a * u_idx.size)      # terms of size a*nnz cancel in the item loss

    # item half-step, WMF flavour (wmf.py:78-96): ridge only, unrated items untouched
    ref_v = fie.copy()
    loss_ref = als_ref.item_step(ref_u, ref_v, i_ptr, i_idx, u_rated, a, b, 0.01, None)
    Vw = V.clone()
    lr = topkrec.als_solve_rows(its, U0, Vw, XXv, a, b, 0.01, 0.01, item_loss=True)
    tol = bound(ref_v, exact64_step(fie, ref_u, i_ptr, i_idx, u_rated, a, b, 0.0, 0.01)) if init == "uniform" or d >= 100 else TOL
    assert rel(Vw.cpu().numpy(), ref_v.astype(np.float64)) <= tol
    assert abs(float(lr.sum()) - loss_ref) <= 1e-5 * (abs(loss_ref) + a * u_idx.size)      # terms of size a*nnz cancel in the item loss


@pytest.mark.parametrize("d,seg", [(12, 16), (64, 33), (100, 4096), (128, 100), (192, 64), (200, 4096), (256, 48)])
@pytest.mark.parametrize("mode", [1, 2])
def test_both_factorisations_at_every_width(d, seg, mode):
    """the blocked 16-column rounds (default at d > 192) and the per-column / four-column loops (default below) are both held to
    the oracle at every width: tkr_debug_set_als_factor 1 = blocked everywhere, 2 = the loops everywhere"""
    L = topkrec.lib()
    L.tkr_debug_set_als_factor(mode)
    try:
        test_user_and_item_steps_match_oracle(d, seg, "uniform")
    finally:
        L.tkr_debug_set_als_factor(0)


def test_split_rows_are_deterministic_and_agree_with_fused():
    rng = np.random.default_rng(5)
    d = 96
    ptr, idx = random_csr(rng, 64, 500, 400)
    Y = torch.from_numpy(rng.random((500, d)).astype(np.float32)).cuda()
    base = topkrec.als_gram(Y, torch.arange(500, dtype=torch.int32, device="cuda"), 0.01, 0.01)
    outs = []
    for seg in (4096, 64, 64, 17):
        X = torch.zeros(64, d, device="cuda")
        topkrec.als_solve_rows(topkrec.AlsSide(ptr, idx, seg), Y, X, base, 1.0, 0.01, 0.0, 0.01)
        outs.append(X.cpu().numpy())
    assert np.array_equal(outs[1], outs[2])
    assert rel(outs[1], outs[0].astype(np.float64)) <= TOL and rel(outs[3], outs[0].astype(np.float64)) <= TOL   # summation order only


def _golden_model(golden, iters):
    from single import CER
    g = np.load(os.path.join(golden, "als_cer.npz"))
    m = CER(k=int(g["k"]), d=int(g["d_feat"]), lu=float(g["lu"]), lv=float(g["lv"]), le=float(g["le"]), a=float(g["a"]), b=float(g["b"]), seg=16)
    n_users, n_items = g["fue0"].shape[0], g["fie0"].shape[0]
    users = np.repeat(np.arange(n_users), g["u_cnt"])
    m.k = int(g["k"])
    m.set_training_pairs(users, g["u_idx"], n_users, n_items, lists=True)
    m.fue, m.fie, m.E, m.feat = g["fue0"].copy(), g["fie0"].copy(), g["E0"].copy(), g["feat"].copy()
    m.train(max_iter=iters, tol=0.0)
    return m, g


@pytest.mark.parametrize("iters", [1, 4])
def test_cer_train_matches_reference_golden(golden, iters):
    """CER.train here vs the UNMODIFIED reference CER.train (tests/golden/make_golden_als.py)."""
    m, g = _golden_model(golden, iters)
    assert rel(m.fue, g["fue%d" % iters].astype(np.float64)) <= TOL
    assert rel(m.fie, g["fie%d" % iters].astype(np.float64)) <= TOL
    assert rel(m.E, g["E%d" % iters]) <= TOL
    assert np.allclose(m.losses, g["losses"][:iters], rtol=2e-5)
    assert m.E.dtype == np.float64 and m.fue.dtype == np.float32            # cer.py:64 leaves E in fp64


def test_wmf_train_matches_oracle(mini):
    from single import WMF
    np.random.seed(3)
    m = WMF(k=20, seg=32)
    with pytest.raises(KeyError):                                           # wmf.py:51: unknown uid in a positive pair
        m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        tr = os.path.join(td, "tr.txt")
        open(tr, "w").write("".join(ln for ln in open(os.path.join(mini, "f0tr.txt")) if not ln.startswith("99999,")))
        m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), tr)
    u_ptr, u_idx, i_ptr, i_idx = m._csr
    assert m.usm[int(m.u_rated[0])] == u_idx[u_ptr[m.u_rated[0]]:u_ptr[m.u_rated[0] + 1]].tolist()
    fue0, fie0 = m.fue.copy(), m.fie.copy()
    ru, rv, rl = als_ref.wmf_train(fue0, fie0, u_ptr, u_idx, i_ptr, i_idx, max_iter=3, tol=0.0)
    m.train(max_iter=3, tol=0.0)
    assert rel(m.fue, ru.astype(np.float64)) <= TOL and rel(m.fie, rv.astype(np.float64)) <= TOL
    assert np.allclose(m.losses, rl, rtol=2e-5)


def test_full_width_residual_property():
    """Size-independent property at d=256 with long rows: the returned x satisfies the normal equations,
    |A x - rhs| / |rhs| small, with A and rhs formed in fp64 on the device."""
    rng = np.random.default_rng(9)
    d, n_y, n_rows = 256, 20000, 300
    ptr, idx = random_csr(rng, n_rows, n_y, 600, empty_frac=0.0, long_rows=((0, 9000), (1, 5000)))
    Y = torch.from_numpy(rng.random((n_y, d)).astype(np.float32)).cuda()
    base = topkrec.als_gram(Y, torch.arange(n_y, dtype=torch.int32, device="cuda"), 0.01, 0.01)
    X = torch.zeros(n_rows, d, device="cuda")
    topkrec.als_solve_rows(topkrec.AlsSide(ptr, idx, 2048), Y, X, base, 1.0, 0.01, 0.0, 0.01)
    Y64, B64 = Y.double(), base.double()
    worst = 0.0
    for r in (0, 1, 2, 150, 299):
        Yi = Y64[torch.from_numpy(idx[ptr[r]:ptr[r + 1]].astype(np.int64)).cuda()]
        A = B64 + 0.99 * (Yi.T @ Yi)
        rhs = Yi.sum(0)
        res = (A @ X[r].double() - rhs).norm() / rhs.norm()
        worst = max(worst, float(res))
    assert worst <= 2e-5, worst


def test_argument_errors():
    Y = torch.zeros(4, 300, device="cuda")
    with pytest.raises(topkrec.TkrError):
        topkrec.als_gram(Y, torch.zeros(1, dtype=torch.int32, device="cuda"), 1.0, 0.0)
    side = topkrec.AlsSide(np.array([0, 1]), np.array([9], np.int32))
    with pytest.raises(ValueError):
        topkrec.als_solve_rows(side, torch.zeros(4, 8, device="cuda"), torch.zeros(1, 8, device="cuda"), torch.zeros(8, 8, device="cuda"), 1, 0.01, 0, 0)


def test_sharded_engine_world1_equals_direct_calls():
    """topkrec.dist.ShardedAls on one rank = the plain calls (bit-identical); the exchange itself is covered by the
    gloo world-2 test in tests/test_dist_cpu.py."""
    from topkrec import dist as tdist
    rng = np.random.default_rng(12)
    nu, ni, d = 300, 90, 40
    u_ptr, u_idx = random_csr(rng, nu, ni, 30)
    users = np.repeat(np.arange(nu), np.diff(u_ptr))
    by_i = np.argsort(u_idx, kind="stable")
    i_ptr = np.zeros(ni + 1, np.int64); np.cumsum(np.bincount(u_idx, minlength=ni), out=i_ptr[1:])
    i_idx = users[by_i].astype(np.int32)
    U0 = torch.from_numpy(rng.random((nu, d)).astype(np.float32)).cuda(); V0 = torch.from_numpy(rng.random((ni, d)).astype(np.float32)).cuda()
    U, V = U0.clone(), V0.clone()
    lu_, li_ = tdist.ShardedAls(u_ptr, u_idx, i_ptr, i_idx, seg=16).iteration(U, V, 1.0, 0.01, 0.01, 0.01, wmf=True)
    us, its = topkrec.AlsSide(u_ptr, u_idx, 16), topkrec.AlsSide(i_ptr, i_idx, 16)
    U2, V2 = U0.clone(), V0.clone()
    l1 = topkrec.als_solve_rows(us, V2, U2, topkrec.als_gram(V2, its.rated_dev, 0.01, 0.01), 1.0, 0.01, 0.0, 0.01)
    l2 = topkrec.als_solve_rows(its, U2, V2, topkrec.als_gram(U2, us.rated_dev, 0.01, 0.0), 1.0, 0.01, 0.01, 0.01, item_loss=True)
    assert torch.equal(U, U2) and torch.equal(V, V2)
    assert abs(lu_ - float(l1.sum())) <= 1e-9 * abs(lu_) and abs(li_ - float(l2.sum())) <= 1e-9 * abs(li_)


def test_degenerate_shapes():
    """no positives at all, a single row, d = 1, one-positive segments: the edge cases of wmf.py:35-56 inputs"""
    rng = np.random.default_rng(21)
    # (a) every row empty: users keep their rows (cer.py:40), CER items are solved from the prior alone (cer.py:61-62)
    d, n = 24, 7
    Y = torch.from_numpy(rng.random((5, d)).astype(np.float32)).cuda()
    X0 = rng.random((n, d)).astype(np.float32)
    side = topkrec.AlsSide(np.zeros(n + 1, np.int64), np.zeros(0, np.int32), 8)
    base = topkrec.als_gram(Y, torch.zeros(0, dtype=torch.int32, device="cuda"), 0.01, 0.5)
    assert np.array_equal(base.cpu().numpy(), 0.5 * np.eye(d, dtype=np.float32))
    X = torch.from_numpy(X0.copy()).cuda()
    loss = topkrec.als_solve_rows(side, Y, X, base, 1.0, 0.01, 0.0, 0.5)
    assert np.array_equal(X.cpu().numpy(), X0)
    assert np.allclose(loss.cpu().numpy(), 0.25 * (X0.astype(np.float64) ** 2).sum(1), rtol=1e-6)
    prior = rng.standard_normal((n, d)).astype(np.float32)
    topkrec.als_solve_rows(side, Y, X, base, 1.0, 0.01, 10.0, 10.0, prior=torch.from_numpy(prior).cuda(), solve_empty=True, item_loss=True)
    assert rel(X.cpu().numpy(), (10.0 / 10.5) * prior.astype(np.float64)) <= 1e-6       # (0.5 I + 10 I) x = 10 prior
    # (b) one row, d = 1, segments of one positive each
    Y1 = torch.from_numpy(np.array([[2.0], [3.0], [0.5]], np.float32)).cuda()
    side1 = topkrec.AlsSide(np.array([0, 3]), np.array([0, 1, 1], np.int32), 1)
    assert side1.n_slots == 3
    X1 = torch.zeros(1, 1, device="cuda")
    b1 = topkrec.als_gram(Y1, torch.tensor([0, 1, 2], dtype=torch.int32, device="cuda"), 0.01, 0.01)
    topkrec.als_solve_rows(side1, Y1, X1, b1, 1.0, 0.01, 0.0, 0.01)
    A = 0.01 * (4 + 9 + 0.25) + 0.01 + 0.99 * (4 + 9 + 9)
    assert abs(float(X1[0, 0]) - (2 + 3 + 3) / A) <= 1e-6

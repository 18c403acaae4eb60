"""GPU parity tests of hot path 1 (tkr_bpr_step / tkr_bpr_sample through the C ABI)
against the CPU oracle.  Tolerance per BASELINE.json north_star: learned U/V within
1e-4 relative (max-norm) after a fixed triple stream and step count; integer work
(the sampler) bit-exact."""
import os

import numpy as np
import pytest
import torch

import topkrec
from oracle import bpr_ref, philox_ref, sampler_ref

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4        # north_star: "learned U/V within 1e-4 relative"


def _to_dev(st):
    return {k: torch.from_numpy(v.copy()).cuda() for k, v in st.items()}


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _run_case(nu, ni, d, B, steps, seed, cfg_kw=None, init_scale=1.0, item_skew=False):
    cfg_kw = cfg_kw or {}
    rng = np.random.default_rng(seed)
    st = bpr_ref.new_state(nu, ni, d, rng)
    st["U"] *= np.float32(init_scale); st["V"] *= np.float32(init_scale)
    st["b"] = (0.01 * init_scale * rng.standard_normal(ni)).astype(np.float32)
    n = B * steps
    u = rng.integers(0, nu, n).astype(np.int32)
    if item_skew:   # popular items: heavy duplicate rows inside a batch (76 of 512 on fold 0)
        p = 1.0 / np.arange(1, ni + 1); p /= p.sum()
        i = rng.choice(ni, n, p=p).astype(np.int32)
    else:
        i = rng.integers(0, ni, n).astype(np.int32)
    j = rng.integers(0, ni, n).astype(np.int32)
    ocfg = bpr_ref.BprCfg(**cfg_kw)
    dst = _to_dev(st)
    cfg = topkrec.BprCfg(nu, ni, d, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lr, ocfg.mode, ocfg.optimizer)
    ws = topkrec.bpr_workspace(cfg, B)
    loss = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.bpr_step(cfg, dst["U"], dst["V"], dst["b"], dst["msU"], dst["msV"], dst["msb"],
                     torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, loss)
    ref_loss = bpr_ref.bpr_train(st, u, i, j, B, ocfg)
    torch.cuda.synchronize()
    for name in st:
        assert _rel(dst[name].cpu().numpy(), st[name]) <= REL_TOL, (name, _rel(dst[name].cpu().numpy(), st[name]))
    assert np.allclose(loss.cpu().numpy(), ref_loss, rtol=1e-4, atol=1e-5)
    _assert_ws_clean(cfg, B, ws)
    return dst, st


def _assert_ws_clean(cfg, B, ws):
    """every accumulator / counter region is zero again after a step (the row lists may hold stale ids)"""
    lay = topkrec.bpr_workspace_layout(cfg, B)
    order = sorted(lay.items(), key=lambda kv: kv[1])
    for (name, beg), (_, end) in zip(order[:-1], order[1:]):
        if name not in ("listU", "listV", "hotV", "stage"):
            assert int(ws[beg:end].count_nonzero().item()) == 0, "workspace region %s is not zero after the step" % name


@pytest.mark.parametrize("d", [128, 50, 33, 64, 200, 256, 512, 7])
def test_bpr_step_matches_oracle_over_widths(d):
    _run_case(300, 200, d, 256, 12, seed=d)


def test_bpr_step_reference_config_d50_b256():
    """C1 shape of the reference recipe (train.py:3-6): k=50, batch 256, default lambdas/lr."""
    _run_case(3000, 800, 50, 256, 100, seed=1, item_skew=True)


def test_bpr_step_heavy_duplicates():
    """tiny tables: every row is hit many times per batch -> the summed-duplicate path dominates."""
    _run_case(7, 5, 128, 512, 20, seed=2, init_scale=20.0)


def test_bpr_step_large_batch():
    _run_case(5000, 1000, 128, 1 << 15, 4, seed=3, item_skew=True)


@pytest.mark.parametrize("kw", [dict(mode="l1"), dict(optimizer="sgd"), dict(optimizer="sgd", mode="l1"),
                                dict(lambda_b=0.05), dict(lr=1e-2, lambda_u=0.1, lambda_i=0.05, lambda_j=0.01, lambda_b=0.02)])
def test_bpr_step_modes(kw):
    _run_case(200, 150, 64, 256, 25, seed=4, cfg_kw=kw, init_scale=10.0)


@pytest.fixture
def count_mode():
    """Force tkr_bpr_step onto the counting path (rows that occur once in a batch are updated in place by the
    gradient kernel; chosen automatically only for tables far beyond L2)."""
    import ctypes
    L = topkrec.lib()
    L.tkr_debug_set_count_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_count_mode.restype = None
    L.tkr_debug_set_count_mode(1)
    yield
    L.tkr_debug_set_count_mode(-1)


@pytest.mark.parametrize("shape", [(300, 200, 128, 256, 12), (3000, 800, 50, 256, 40), (7, 5, 128, 512, 8),
                                   (50000, 30000, 128, 4096, 6), (400, 300, 33, 64, 10), (5000, 1000, 256, 1 << 14, 3)])
def test_bpr_step_count_mode_matches_oracle(count_mode, shape):
    """same parity bar on the counting path: mixes of once-only rows (in place) and duplicated rows (accumulators)"""
    nu, ni, d, B, steps = shape
    _run_case(nu, ni, d, B, steps, seed=11 + d, item_skew=nu < 10000)


@pytest.mark.parametrize("kw", [dict(optimizer="sgd"), dict(lambda_b=0.05), dict(lr=1e-2, lambda_u=0.1, lambda_i=0.05, lambda_j=0.01)])
def test_bpr_step_count_mode_options(count_mode, kw):
    _run_case(2000, 1500, 64, 256, 25, seed=12, cfg_kw=kw, init_scale=10.0)


def test_bpr_step_count_mode_equals_accumulator_path():
    """the two paths of tkr_bpr_step agree to rounding on the same stream (same update expression, different
    summation route for the once-only rows: 0 + g)"""
    import ctypes
    L = topkrec.lib()
    L.tkr_debug_set_count_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_count_mode.restype = None
    rng = np.random.default_rng(13)
    nu, ni, d, B, steps = 20000, 9000, 128, 2048, 5
    st = bpr_ref.new_state(nu, ni, d, rng)
    u = torch.from_numpy(rng.integers(0, nu, B * steps).astype(np.int32)).cuda()
    i = torch.from_numpy(rng.integers(0, ni, B * steps).astype(np.int32)).cuda()
    j = torch.from_numpy(rng.integers(0, ni, B * steps).astype(np.int32)).cuda()
    cfg = topkrec.BprCfg(nu, ni, d)
    outs = []
    for mode in (0, 1):
        L.tkr_debug_set_count_mode(mode)
        dst = _to_dev(st)
        ws = topkrec.bpr_workspace(cfg, B)
        topkrec.bpr_step(cfg, dst["U"], dst["V"], dst["b"], dst["msU"], dst["msV"], dst["msb"], u, i, j, B, steps, ws)
        outs.append({k: v.cpu().numpy() for k, v in dst.items()})
        _assert_ws_clean(cfg, B, ws)
    L.tkr_debug_set_count_mode(-1)
    for name in outs[0]:
        assert _rel(outs[1][name], outs[0][name]) <= 1e-6, name


def _set_persist(mode):
    import ctypes
    L = topkrec.lib()
    L.tkr_debug_set_persist_mode.argtypes = [ctypes.c_int32]; L.tkr_debug_set_persist_mode.restype = None
    L.tkr_debug_set_persist_mode(mode)


@pytest.fixture
def two_kernel_route():
    """tkr_bpr_step takes the persistent cluster kernel for batches <= 1024 by default; this forces the two-launch route"""
    _set_persist(0)
    yield
    _set_persist(-1)


@pytest.mark.parametrize("shape", [(300, 200, 128, 256, 12), (3000, 800, 50, 256, 100), (7, 5, 128, 512, 20), (400, 300, 33, 64, 10),
                                   (900, 700, 256, 1024, 6), (2000, 1500, 200, 700, 7), (50, 40, 7, 1, 30)])
def test_bpr_step_two_kernel_route_matches_oracle(two_kernel_route, shape):
    """the non-persistent route for small batches keeps the same parity bar (it is what larger batches and wide rows use)"""
    nu, ni, d, B, steps = shape
    _run_case(nu, ni, d, B, steps, seed=31 + d, item_skew=nu >= 300)


@pytest.fixture
def persistent_route():
    """force the persistent cluster kernel wherever it is legal (B <= 1024; the automatic choice stops at B = 256)"""
    _set_persist(1)
    yield
    _set_persist(-1)


@pytest.fixture
def dataflow_route():
    """force the dataflow multi-step kernel (tkr_debug_set_persist_mode 2): no grid-wide barriers, every triple waits only for
    the rows it reads (version words), the last occurrence of a row in a step applies its update"""
    _set_persist(2)
    yield
    _set_persist(-1)


@pytest.mark.parametrize("shape", [(3000, 800, 50, 256, 100), (900, 700, 256, 1024, 6), (2000, 1500, 200, 700, 7), (50, 40, 7, 1, 30), (300, 200, 128, 97, 9),
                                   (300, 200, 128, 256, 300), (40, 12, 64, 256, 40), (70000, 10000, 128, 256, 64)])
def test_bpr_step_dataflow_kernel_matches_oracle(dataflow_route, shape):
    """same parity bar as the other two routes: popular items (a row touched in every step: the serial chain of the version
    words), tiny tables (every row occurs many times per step, i == j triples), chunks of more than 256 steps (two launches),
    ragged batch sizes, a single triple, the C2 table sizes"""
    nu, ni, d, B, steps = shape
    _run_case(nu, ni, d, B, steps, seed=57 + d + B, item_skew=ni >= 200)


@pytest.mark.parametrize("shape", [(3000, 800, 50, 256, 100), (900, 700, 256, 1024, 6), (2000, 1500, 200, 700, 7), (50, 40, 7, 1, 30), (300, 200, 128, 97, 9)])
def test_bpr_step_persistent_kernel_matches_oracle(persistent_route, shape):
    """persistent cluster kernel (default for B <= 1024, d <= 256): one triple per warp (B <= 256) and the list path
    (256 < B <= 1024), ragged batch sizes, a single triple"""
    nu, ni, d, B, steps = shape
    _run_case(nu, ni, d, B, steps, seed=41 + d, item_skew=nu >= 300)


@pytest.mark.parametrize("kw", [dict(mode="l1"), dict(optimizer="sgd"), dict(lambda_b=0.05), dict(lr=1e-2, lambda_u=0.1, lambda_i=0.05, lambda_j=0.01, lambda_b=0.02)])
def test_bpr_step_persistent_kernel_modes(persistent_route, kw):
    _run_case(200, 150, 64, 256, 25, seed=42, cfg_kw=kw, init_scale=10.0)
    _run_case(200, 150, 64, 600, 8, seed=43, cfg_kw=kw, init_scale=10.0)


def test_bpr_step_persistent_equals_two_kernel_route_and_repeated_rows():
    """same stream through both routes: equal to atomic-order noise; includes triples with i == j and a batch made of one
    user (every warp claims the same row: exactly one updater)"""
    rng = np.random.default_rng(44)
    nu, ni, d, B, steps = 500, 300, 128, 256, 30
    st = bpr_ref.new_state(nu, ni, d, rng)
    u = rng.integers(0, nu, B * steps).astype(np.int32); i = rng.integers(0, ni, B * steps).astype(np.int32); j = rng.integers(0, ni, B * steps).astype(np.int32)
    j[:40] = i[:40]                      # degenerate triples (explicit streams may contain them)
    u[B:2 * B] = 7                       # step 1: one user, 256 occurrences
    i[2 * B:3 * B] = 11                  # step 2: one positive item
    cfg = topkrec.BprCfg(nu, ni, d, lambda_b=0.01)
    outs = []
    for mode in (0, 1):
        _set_persist(mode)
        dst = _to_dev(st)
        ws = topkrec.bpr_workspace(cfg, B)
        loss = torch.empty(steps, device="cuda")
        topkrec.bpr_step(cfg, dst["U"], dst["V"], dst["b"], dst["msU"], dst["msV"], dst["msb"], torch.from_numpy(u).cuda(),
                         torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, loss)
        outs.append({k: v.cpu().numpy() for k, v in dst.items()} | {"loss": loss.cpu().numpy()})
        _assert_ws_clean(cfg, B, ws)
    _set_persist(-1)
    for name in outs[0]:
        assert _rel(outs[1][name], outs[0][name]) <= 2e-6, name
    ref = {k: v.copy() for k, v in st.items()}
    ref_loss = bpr_ref.bpr_train(ref, u, i, j, B, bpr_ref.BprCfg(lambda_b=0.01))
    for name in ref:
        assert _rel(outs[1][name], ref[name]) <= REL_TOL, name
    assert np.allclose(outs[1]["loss"], ref_loss, rtol=1e-4)


def test_bpr_step_persistent_fused_sampler_chunks(mini):
    """fused sampler + persistent kernel: the triples of a chunk of steps are drawn into the workspace first; 700 steps of
    256 cross the 65 536-triple staging buffer twice and must equal sampling everything up front"""
    tr_users, tr_data, indptr, idx, nu, ni = _mini_tables(mini)
    rng = np.random.default_rng(45)
    d, B, steps, first = 64, 256, 700, 12345
    st = bpr_ref.new_state(nu, ni, d, rng)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=77)
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    a, b2 = _to_dev(st), _to_dev(st)
    la = torch.empty(steps, device="cuda"); lb = torch.empty(steps, device="cuda")
    topkrec.bpr_step(cfg, a["U"], a["V"], a["b"], a["msU"], a["msV"], a["msb"], None, None, None, B, steps, ws, la, sampler=smp, first_draw=first)
    u, i, j = topkrec.bpr_sample(smp, first, B * steps)
    topkrec.bpr_step(cfg, b2["U"], b2["V"], b2["b"], b2["msU"], b2["msV"], b2["msb"], u, i, j, B, steps, ws, lb)
    for n in a:
        assert _rel(a[n].cpu().numpy(), b2[n].cpu().numpy()) <= 2e-6, n
    assert np.allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=1e-5)
    _assert_ws_clean(cfg, B, ws)


@pytest.mark.parametrize("B,steps", [(48, 3000), (512, 300)])
def test_bpr_step_dataflow_fused_sampler_chunks(mini, B, steps):
    """fused sampler + dataflow kernel (the automatic choice at these batch sizes): the staged draws cross the 65 536-triple
    buffer twice, every chunk gets its own pre-pass and launch, the version words carry over; equal to sampling up front (which
    takes the 2^18-triple chunks of explicit triples) and to the two-launch route"""
    tr_users, tr_data, indptr, idx, nu, ni = _mini_tables(mini)
    rng = np.random.default_rng(46)
    d, first = 64, 4321
    st = bpr_ref.new_state(nu, ni, d, rng)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=78)
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    a, b2, c = _to_dev(st), _to_dev(st), _to_dev(st)
    la = torch.empty(steps, device="cuda"); lb = torch.empty(steps, device="cuda"); lc = torch.empty(steps, device="cuda")
    topkrec.bpr_step(cfg, a["U"], a["V"], a["b"], a["msU"], a["msV"], a["msb"], None, None, None, B, steps, ws, la, sampler=smp, first_draw=first)
    u, i, j = topkrec.bpr_sample(smp, first, B * steps)
    topkrec.bpr_step(cfg, b2["U"], b2["V"], b2["b"], b2["msU"], b2["msV"], b2["msb"], u, i, j, B, steps, ws, lb)
    _set_persist(0)
    try:
        topkrec.bpr_step(cfg, c["U"], c["V"], c["b"], c["msU"], c["msV"], c["msb"], u, i, j, B, steps, ws, lc)
    finally:
        _set_persist(-1)
    for n in a:
        assert _rel(a[n].cpu().numpy(), b2[n].cpu().numpy()) <= 2e-6, n
        assert _rel(a[n].cpu().numpy(), c[n].cpu().numpy()) <= 2 * REL_TOL, n      # (two routes, each within REL_TOL of the oracle)
    assert np.allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=1e-5)
    assert np.allclose(la.cpu().numpy(), lc.cpu().numpy(), rtol=1e-4)
    _assert_ws_clean(cfg, B, ws)


@pytest.mark.parametrize("shape", [(3000, 800, 50, 256, 40), (5000, 1000, 128, 1 << 15, 4), (7, 5, 128, 512, 8), (900, 700, 256, 4096, 5),
                                   (400, 300, 33, 64, 10)])
def test_bpr_step_hot_items_match_oracle(shape):
    """popular item rows privatised per thread block in shared memory: same sums, same parity bar"""
    nu, ni, d, B, steps = shape
    rng = np.random.default_rng(17 + d)
    st = bpr_ref.new_state(nu, ni, d, rng)
    st["b"] = (0.01 * rng.standard_normal(ni)).astype(np.float32)
    n = B * steps
    p = 1.0 / np.arange(1, ni + 1); p /= p.sum()
    perm = rng.permutation(ni)                                   # popular items are not the low ids
    u = rng.integers(0, nu, n).astype(np.int32)
    i = perm[rng.choice(ni, n, p=p)].astype(np.int32)
    j = rng.integers(0, ni, n).astype(np.int32)
    ocfg = bpr_ref.BprCfg(lambda_b=0.01)
    dst = _to_dev(st)
    cfg = topkrec.BprCfg(nu, ni, d, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lr, ocfg.mode, ocfg.optimizer)
    ws = topkrec.bpr_workspace(cfg, B)
    hot = topkrec.popular_items(i, ni)
    assert 0 < hot.size <= topkrec.MAX_HOT
    topkrec.bpr_set_hot_items(cfg, B, ws, hot)
    loss = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.bpr_step(cfg, dst["U"], dst["V"], dst["b"], dst["msU"], dst["msV"], dst["msb"],
                     torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, loss)
    ref_loss = bpr_ref.bpr_train(st, u, i, j, B, ocfg)
    for name in st:
        assert _rel(dst[name].cpu().numpy(), st[name]) <= REL_TOL, (name, _rel(dst[name].cpu().numpy(), st[name]))
    assert np.allclose(loss.cpu().numpy(), ref_loss, rtol=1e-4, atol=1e-5)
    _assert_ws_clean(cfg, B, ws)
    topkrec.bpr_set_hot_items(cfg, B, ws, [])                    # clearing restores an all-zero map
    lay = topkrec.bpr_workspace_layout(cfg, B)
    assert int(ws[lay["hotV"]:lay["hotV"] + 4 * ni].count_nonzero().item()) == 0
    with pytest.raises(topkrec.TkrError):
        topkrec.bpr_set_hot_items(cfg, B, ws, [0, 0])


def test_bpr_step_lazy_rows_and_single_update():
    """Untouched rows and slots keep their bits; a touched row's rms slot moved exactly once."""
    rng = np.random.default_rng(5)
    nu, ni, d, B = 1000, 600, 128, 64
    st = bpr_ref.new_state(nu, ni, d, rng)
    dst = _to_dev(st)
    u = rng.integers(0, 100, B).astype(np.int32); i = rng.integers(0, 50, B).astype(np.int32); j = rng.integers(50, 100, B).astype(np.int32)
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    topkrec.bpr_step(cfg, dst["U"], dst["V"], dst["b"], dst["msU"], dst["msV"], dst["msb"],
                     torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, 1, ws)
    U = dst["U"].cpu().numpy(); msU = dst["msU"].cpu().numpy(); msV = dst["msV"].cpu().numpy()
    touched = np.zeros(nu, bool); touched[u] = True
    assert np.array_equal(U[~touched], st["U"][~touched]) and (msU[~touched] == 1).all()
    assert (msU[touched] < 1).all() and (msU[touched] >= 0.9).all()          # 0.9*1 + 0.1*g^2, g small: one decay
    tv = np.zeros(ni, bool); tv[i] = True; tv[j] = True
    assert (msV[~tv] == 1).all() and (msV[tv] < 1).all() and (msV[tv] >= 0.9).all()


def test_bpr_step_host_entry_equals_device_entry():
    rng = np.random.default_rng(6)
    nu, ni, d, B, steps = 400, 300, 128, 256, 6
    st = bpr_ref.new_state(nu, ni, d, rng)
    a, b = _to_dev(st), _to_dev(st)
    u = rng.integers(0, nu, B * steps).astype(np.int32); i = rng.integers(0, ni, B * steps).astype(np.int32); j = rng.integers(0, ni, B * steps).astype(np.int32)
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    la = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.bpr_step(cfg, a["U"], a["V"], a["b"], a["msU"], a["msV"], a["msb"], torch.from_numpy(u).cuda(),
                     torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, ws, la)
    lb = torch.empty(steps, dtype=torch.float32).pin_memory()
    staging = torch.empty(4 * (B * steps * 4 + 256), dtype=torch.uint8, device="cuda")
    topkrec.bpr_step_host(cfg, b["U"], b["V"], b["b"], b["msU"], b["msV"], b["msb"], torch.from_numpy(u).pin_memory(),
                          torch.from_numpy(i).pin_memory(), torch.from_numpy(j).pin_memory(), B, steps, lb, staging, ws)
    for n in a:
        assert _rel(b[n].cpu().numpy(), a[n].cpu().numpy()) <= 1e-6, n       # same kernels; only atomic order differs
    assert np.allclose(lb.numpy(), la.cpu().numpy(), rtol=1e-5)


def test_data_parallel_halves_equal_one_big_batch():
    """tkr_bpr_grad on two user-partitioned half batches + summed item gradients + dense apply
    == one step over the union batch (what 2 ranks + all-reduce compute)."""
    rng = np.random.default_rng(9)
    nu, ni, d, B = 600, 200, 128, 512
    st = bpr_ref.new_state(nu, ni, d, rng)
    u = rng.integers(0, nu, 2 * B).astype(np.int32); i = rng.integers(0, ni, 2 * B).astype(np.int32); j = rng.integers(0, ni, 2 * B).astype(np.int32)
    order = np.argsort(u % 2, kind="stable")                     # rank r owns users u % 2 == r
    u, i, j = u[order], i[order], j[order]
    nb0 = int((u % 2 == 0).sum())
    parts = [(0, nb0), (nb0, 2 * B)]
    cfg = topkrec.BprCfg(nu, ni, d)
    ranks = []
    for beg, end in parts:
        s = _to_dev(st); ws = topkrec.bpr_workspace(cfg, end - beg); loss = torch.zeros(1, device="cuda")
        topkrec.bpr_grad(cfg, s["U"], s["V"], s["b"], torch.from_numpy(u[beg:end]).cuda(), torch.from_numpy(i[beg:end]).cuda(),
                         torch.from_numpy(j[beg:end]).cuda(), end - beg, ws, loss, data_parallel=True)
        ranks.append((s, ws, loss, end - beg))
    views = [topkrec.bpr_item_grad_view(cfg, n, ws) for _, ws, _, n in ranks]
    total = views[0] + views[1]                                   # the all-reduce
    for v in views:
        v.copy_(total)
    for s, ws, _, n in ranks:
        topkrec.bpr_apply(cfg, s["U"], s["V"], s["b"], s["msU"], s["msV"], s["msb"], n, ws, data_parallel=True)
        _assert_ws_clean(cfg, n, ws)
    ref_loss = bpr_ref.bpr_step(st, u, i, j, bpr_ref.BprCfg())
    assert abs((ranks[0][2] + ranks[1][2]).item() - ref_loss) / ref_loss < 1e-5
    for name in ("V", "b", "msV", "msb"):                         # replicas identical and equal to the big batch
        assert torch.equal(ranks[0][0][name], ranks[1][0][name])
        assert _rel(ranks[0][0][name].cpu().numpy(), st[name]) <= REL_TOL, name
    for name in ("U", "msU"):                                     # each rank updated only its own users
        got = ranks[0][0][name].cpu().numpy().copy()
        got[1::2] = ranks[1][0][name].cpu().numpy()[1::2]
        assert _rel(got, st[name]) <= REL_TOL, name


def test_error_paths():
    cfg = topkrec.BprCfg(10, 10, 8)
    ws = topkrec.bpr_workspace(cfg, 16)
    t = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    z = torch.zeros(16, dtype=torch.int32, device="cuda")
    with pytest.raises(topkrec.TkrError, match="workspace too small"):
        topkrec.bpr_step(cfg, t(10, 8), t(10, 8), t(10), t(10, 8), t(10, 8), t(10), z, z, z, 16, 1, ws[:256])
    with pytest.raises(topkrec.TkrError, match="msU"):
        topkrec.bpr_step(cfg, t(10, 8), t(10, 8), t(10), None, None, None, z, z, z, 16, 1, ws)
    big = topkrec.BprCfg(10, 10, 4096 + 4)
    with pytest.raises(topkrec.TkrError, match="too wide"):
        topkrec.bpr_step(big, t(10, 4100), t(10, 4100), t(10), t(10, 4100), t(10, 4100), t(10), z, z, z, 16, 1,
                         topkrec.bpr_workspace(big, 16))


# ------------------------------------------------------------------ sampler (integer work: bit-exact)
def _mini_tables(mini):
    uids = sampler_ref.load_ids(os.path.join(mini, "uid")); iids = sampler_ref.load_ids(os.path.join(mini, "vid"))
    tr_users, tr_data, _ = sampler_ref.load_positives(os.path.join(mini, "f0tr.txt"), uids, iids)
    indptr, idx = sampler_ref.to_csr(tr_users, tr_data, len(uids))
    for uu in tr_users:
        idx[indptr[uu]:indptr[uu + 1]].sort()
    return tr_users, tr_data, indptr, idx, len(uids), len(iids)


def test_device_sampler_bit_exact(mini):
    tr_users, tr_data, indptr, idx, nu, ni = _mini_tables(mini)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=0xDEADBEEF12345)
    for first, n in ((0, 5000), (123456789012, 777), ((1 << 32) - 100, 300)):
        u, i, j = topkrec.bpr_sample(smp, first, n)
        ru, ri, rj = philox_ref.sample(tr_users, indptr, idx, ni, 0xDEADBEEF12345, first, n)
        assert np.array_equal(u.cpu().numpy(), ru) and np.array_equal(i.cpu().numpy(), ri) and np.array_equal(j.cpu().numpy(), rj)


def test_device_sampler_rejection_heavy():
    """a user who likes all but one item: the redraw loop must land on the single negative."""
    ni = 40
    tr_users = [0, 1]
    indptr = np.array([0, ni - 1, ni + 1], np.int64)
    idx = np.concatenate([np.delete(np.arange(ni), 17), [3, 9]]).astype(np.int32)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=7)
    u, i, j = (t.cpu().numpy() for t in topkrec.bpr_sample(smp, 0, 2000))
    ru, ri, rj = philox_ref.sample(tr_users, indptr, idx, ni, 7, 0, 2000)
    assert np.array_equal(u, ru) and np.array_equal(i, ri) and np.array_equal(j, rj)
    assert (j[u == 0] == 17).all()


def test_fused_sampling_equals_sample_then_step(mini):
    tr_users, tr_data, indptr, idx, nu, ni = _mini_tables(mini)
    rng = np.random.default_rng(8)
    d, B, steps, first = 128, 256, 5, 1000
    st = bpr_ref.new_state(nu, ni, d, rng)
    smp = topkrec.Sampler(tr_users, indptr, idx, ni, seed=42)
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    a = _to_dev(st)
    la = torch.empty(steps, dtype=torch.float32, device="cuda")
    topkrec.bpr_step(cfg, a["U"], a["V"], a["b"], a["msU"], a["msV"], a["msb"], None, None, None, B, steps, ws, la,
                     sampler=smp, first_draw=first)
    u, i, j = philox_ref.sample(tr_users, indptr, idx, ni, 42, first, B * steps)
    ref_loss = bpr_ref.bpr_train(st, u, i, j, B, bpr_ref.BprCfg())
    for n in st:
        assert _rel(a[n].cpu().numpy(), st[n]) <= REL_TOL, n
    assert np.allclose(la.cpu().numpy(), ref_loss, rtol=1e-4)


def test_full_size_properties_c2():
    """BASELINE config 2 shape (70k x 10k, d=128) at B=2^20: size-independent properties --
    loss at N(0,0.01) init is ~ B ln2; lazy rows untouched; one decay per touched row; a second call with
    the same draws from the same state is reproducible to atomic-order noise."""
    nu, ni, d, B = 70000, 10000, 128, 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    mk = lambda: {"U": torch.randn(nu, d, device="cuda", generator=g) * 0.01, "V": torch.randn(ni, d, device="cuda", generator=g) * 0.01,  # noqa: E731
                  "b": torch.zeros(ni, device="cuda")}
    a = mk(); a.update({"ms" + k: torch.ones_like(v) for k, v in list(a.items())})
    b = {k: v.clone() for k, v in a.items()}
    rng = np.random.default_rng(0)
    u = torch.from_numpy(rng.integers(0, nu // 2, B).astype(np.int32)).cuda()       # only half the users are touched
    p = 1.0 / np.arange(1, ni + 1); p /= p.sum()
    i = torch.from_numpy(rng.choice(ni, B, p=p).astype(np.int32)).cuda()
    j = torch.from_numpy(rng.integers(0, ni, B).astype(np.int32)).cuda()
    cfg = topkrec.BprCfg(nu, ni, d)
    ws = topkrec.bpr_workspace(cfg, B)
    U0 = a["U"].clone()
    la = torch.empty(1, device="cuda"); lb = torch.empty(1, device="cuda")
    topkrec.bpr_step(cfg, a["U"], a["V"], a["b"], a["msU"], a["msV"], a["msb"], u, i, j, B, 1, ws, la)
    topkrec.bpr_step(cfg, b["U"], b["V"], b["b"], b["msU"], b["msV"], b["msb"], u, i, j, B, 1, ws, lb)
    assert abs(la.item() / (B * np.log(2)) - 1) < 1e-2
    assert torch.equal(a["U"][nu // 2:], U0[nu // 2:]) and bool((a["msU"][nu // 2:] == 1).all())
    cnt = torch.bincount(u.long(), minlength=nu)
    assert bool((a["msU"][cnt > 0] < 1).all()) and bool((a["msU"][cnt > 0] >= 0.9).all())
    for n in a:   # (item 0 collects ~1e5 gradients of both signs per step: its squared sum carries ~1e-5 of order noise)
        assert _rel(a[n].cpu().numpy(), b[n].cpu().numpy()) <= 5e-5, n
    _assert_ws_clean(cfg, B, ws)


def test_bpr_class_train_export_resume(tmp_path, mini):
    """train.py:3-9 on the mini fixture through the public class: device sampler, export, warm start."""
    import single
    m = single.BPR(k=16, seed=3)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    m.train(epochs=3, batch_size=64, epoch_sample_limit=20e2)          # float limit as in train.py:6 (D-1)
    steps = 3 * (2000 // 64)
    assert len(m.losses) == steps and m.losses[-1] < m.losses[0] * 1.05
    assert m.fue.shape == (300, 16) and m.fie.shape == (160, 16) and m.fib.shape == (160, 1) and m.fue.dtype == np.float32
    out = str(tmp_path / "embed" / "bpr")
    m.export_embeddings(out)
    assert sorted(os.listdir(out)) == ["final-B.dat", "final-U.dat", "final-V.dat", "weights.npz"]
    fue = m.fue.copy()
    m.train(epochs=1, batch_size=64, epoch_sample_limit=64, model_path=out)   # one step from the exported model
    assert np.abs(m.fue - fue).max() < 1e-3                                    # started from the export, moved a little


def test_bpr_class_numpy_sampler_matches_oracle_replay(mini):
    """sampler='numpy' replays the reference RNG stream: same triples as oracle.sampler_ref -> state within 1e-4."""
    import single
    rng = np.random.default_rng(11)
    m = single.BPR(k=32, sampler="numpy", seed=1)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    st = bpr_ref.new_state(m.n_users, m.n_items, 32, rng)
    m.fue, m.fie, m.fib = st["U"].copy(), st["V"].copy(), st["b"].reshape(-1, 1).copy()   # inject initial weights (bpr.py:127-135)
    np.random.seed(123)
    m.train(epochs=1, batch_size=64, epoch_sample_limit=64 * 20)
    rs = np.random.RandomState(123)
    ub, ib, jb = sampler_ref.replay_sampler(m.tr_users, m.tr_data, m.n_items, 64, 20, rs)
    ref_loss = bpr_ref.bpr_train(st, ub.ravel(), ib.ravel(), jb.ravel(), 64, bpr_ref.BprCfg())
    assert _rel(m.fue, st["U"]) <= REL_TOL and _rel(m.fie, st["V"]) <= REL_TOL and _rel(m.fib.ravel(), st["b"]) <= REL_TOL
    assert np.allclose(m.losses, ref_loss, rtol=1e-4)


def test_hogwild_sgd_equals_synchronous_sgd_without_row_conflicts_and_trains():
    """tkr_bpr_hogwild (SURVEY 8(f) NEXT-4): with no row occurring twice in a batch the barrier-free update IS the synchronous
    SGD step (old/methods/bpr.py:57-61) -- checked against the oracle; with conflicts it is only required to train."""
    rng = np.random.default_rng(61)
    nu, ni, d, B, steps = 4000, 9000, 128, 1024, 6
    st = bpr_ref.new_state(nu, ni, d, rng)
    st["U"] *= 20; st["V"] *= 20
    u = np.concatenate([rng.permutation(nu)[:B] for _ in range(steps)]).astype(np.int32)
    ij = np.stack([rng.permutation(ni)[:2 * B] for _ in range(steps)])
    i, j = ij[:, :B].ravel().astype(np.int32), ij[:, B:].ravel().astype(np.int32)          # every row at most once per batch
    ocfg = bpr_ref.BprCfg(optimizer="sgd", lr=0.05, lambda_b=0.01)
    cfg = topkrec.BprCfg(nu, ni, d, ocfg.lambda_u, ocfg.lambda_i, ocfg.lambda_j, ocfg.lambda_b, ocfg.lr, "l2", "sgd")
    dst = _to_dev(st)
    loss = torch.empty(steps, device="cuda")
    topkrec.bpr_hogwild(cfg, dst["U"], dst["V"], dst["b"], torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda(), B, steps, loss)
    ref_loss = bpr_ref.bpr_train(st, u, i, j, B, ocfg)
    for n in ("U", "V", "b"):
        assert _rel(dst[n].cpu().numpy(), st[n]) <= REL_TOL, n
    assert np.allclose(loss.cpu().numpy(), ref_loss, rtol=1e-4)
    # heavy conflicts (a few popular rows): still descends
    n_tr = 4096 * 60
    u2 = rng.integers(0, 50, n_tr).astype(np.int32); i2 = rng.choice(200, n_tr).astype(np.int32); j2 = (200 + rng.integers(0, 200, n_tr)).astype(np.int32)
    l2 = torch.empty(60, device="cuda")
    topkrec.bpr_hogwild(cfg, dst["U"], dst["V"], dst["b"], torch.from_numpy(u2).cuda(), torch.from_numpy(i2).cuda(), torch.from_numpy(j2).cuda(), 4096, 60, l2)
    l2 = l2.cpu().numpy()
    assert np.isfinite(l2).all() and l2[-1] < 0.7 * l2[0]


def test_bpr_class_hogwild_sampling_switch(mini):
    import single
    m = single.BPR(k=16, seed=3, lr=0.05)
    m.load_training_data(os.path.join(mini, "uid"), os.path.join(mini, "vid"), os.path.join(mini, "f0tr.txt"))
    m.train(sampling="user uniform hogwild", epochs=4, batch_size=64, epoch_sample_limit=2000)
    assert np.isfinite(m.fue).all() and m.losses[-1] < m.losses[0]

"""GPU parity tests of hot path 2 (tkr_score_topk / tkr_topk_merge through the C ABI):
index lists and scores must be BIT-EXACT against the C oracle (oracle/topk_ref.c), which is
itself pinned to the reference's evaluate.py by the committed goldens."""
import json
import os

import numpy as np
import pytest
import torch

import topkrec
from oracle import topk_ref

pytestmark = pytest.mark.gpu


def _case(nu, ni, d, k, seed, bias=False, rated=0, ties=False, scale=0.1):
    rng = np.random.default_rng(seed)
    U = (scale * rng.standard_normal((nu, d))).astype(np.float32)
    V = (scale * rng.standard_normal((ni, d))).astype(np.float32)
    if ties:
        V[rng.integers(0, ni, ni // 10)] = V[0]
        V[rng.integers(0, ni, ni // 20)] = 0
        U[nu // 2] = 0
    b = (0.05 * rng.standard_normal(ni)).astype(np.float32) if bias else None
    indptr = idx = None
    if rated:
        cnt = rng.integers(0, rated + 1, nu)
        cnt[0] = 0
        if nu > 3:
            cnt[3] = min(ni, 4 * rated)
        indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum(cnt)
        idx = np.concatenate([np.sort(rng.choice(ni, c, replace=False)) for c in cnt] + [np.zeros(0, np.int64)]).astype(np.int32)
    return U, V, b, indptr, idx


ENGINES = ("exact", "tc")     # CUDA-core fp32 kernel / tcgen05 BF16 filter + exact refine: same bits required of both


def _check(U, V, k, b, indptr, idx, col_offset=0, engines=ENGINES):
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()  # noqa: E731
    ri, rs = topk_ref.score_topk(U, V, k, b, indptr, idx, col_offset=col_offset)
    fallback = {}
    for eng in engines:
        nf = torch.zeros(1, dtype=torch.int32, device="cuda")
        gi, gs = topkrec.score_topk(t(U), t(V), k, t(b), t(indptr), t(idx), col_offset=col_offset, engine=eng, n_fallback=nf)
        gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
        assert np.array_equal(gi, ri), "[%s] index lists differ at rows %s" % (eng, np.nonzero((gi != ri).any(1))[0][:5])
        assert np.array_equal(gs.view(np.uint32), rs.view(np.uint32)), "[%s] scores are not bit-identical" % eng
        fallback[eng] = int(nf.item())
    return fallback


@pytest.mark.parametrize("nu,ni,d,k", [(300, 1000, 128, 30), (129, 65, 50, 30), (1, 1, 1, 1), (77, 200, 16, 5),
                                       (64, 333, 256, 30), (40, 500, 300, 30), (33, 700, 512, 10), (10, 90, 7, 64),
                                       (500, 4099, 64, 30), (260, 2048, 128, 15)])
def test_score_topk_bit_exact_shapes(nu, ni, d, k):
    fb = _check(*_case(nu, ni, d, k, seed=nu + ni)[:2], k, None, None, None)
    if ni >= 1000 and d <= 256:
        assert fb["tc"] <= nu // 50, "tensor-core filter should certify nearly every row on generic data (%d fell back)" % fb["tc"]


@pytest.mark.parametrize("nu,ni,d,bias,rated", [
    (300, 70000, 128, False, 0),      # 274 tiles: seeded sweep (22 seed tiles spread with stride 12), second CTA pair half empty
    (700, 131072 + 77, 150, True, 50),   # d + 3 = 153 -> three K chunks, padded to four; ragged last tile; bias columns; mask
    (260, 200000, 64, False, 300),    # one K chunk per tile (8-stage ring), long rated lists (filtered at compaction)
    (40, 300000, 250, True, 0),       # few users: item splits over the chip (clusters along y), d + 3 = 253
    (513, 66000, 96, False, 20),      # 5 CTA pairs, seeded, two K chunks with a half-empty second one
])
def test_score_topk_tc_pipeline_shapes(nu, ni, d, bias, rated):
    """shapes that exercise the CTA-pair filter's corner paths (tensor-core engine only: the exact engine on these
    sizes is slow and is itself pinned by the small cases); the oracle is the reference of both"""
    U, V, b, p, i = _case(nu, ni, d, 30, seed=nu + d, bias=bias, rated=rated)
    fb = _check(U, V, 30, b, p, i, engines=("tc",))
    assert fb["tc"] <= max(2, nu // 50), "%d rows fell back to the exact engine" % fb["tc"]


def test_score_topk_tc_biased_item_order():
    """items sorted by norm (the best ones first): the seed tiles are spread over the sweep, so the seed threshold is
    an unbiased sample and the rows still certify"""
    rng = np.random.default_rng(31)
    nu, ni, d = 256, 80000, 128
    U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
    V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
    V *= np.linspace(3.0, 0.3, ni, dtype=np.float32)[:, None]
    fb = _check(U, V, 30, None, None, None, engines=("tc",))
    assert fb["tc"] <= 8


def test_score_topk_mostly_rated_top():
    """every user has rated nearly all of its best columns: the candidate buffers fill with rated columns between
    compactions and the filter must still deliver the first 30 unrated ones"""
    rng = np.random.default_rng(32)
    nu, ni, d, k = 200, 30000, 64, 30
    U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
    V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
    S = U @ V.T
    top = np.argsort(-S, axis=1)[:, :400]
    rated = [np.sort(rng.choice(top[r], 380, replace=False)) for r in range(nu)]
    indptr = np.zeros(nu + 1, np.int64); indptr[1:] = np.cumsum([len(x) for x in rated])
    idx = np.concatenate(rated).astype(np.int32)
    _check(U, V, k, None, indptr, idx)


@pytest.mark.parametrize("k", [5, 10, 15, 20, 25, 30])
def test_score_topk_reference_cutoffs(k):
    """the reference's k in {5,...,30} (evaluate.py -s 5 -t 30)"""
    U, V, b, p, i = _case(200, 1500, 128, k, seed=k, bias=True, rated=40, ties=True)
    _check(U, V, k, b, p, i)


def test_score_topk_ties_and_zero_rows():
    U, V, b, p, i = _case(150, 900, 50, 30, seed=11, ties=True)
    _check(U, V, 30, None, None, None)
    V[:] = 0                                   # every score ties at +0: order = column descending
    t = lambda a: torch.from_numpy(a).cuda()   # noqa: E731
    for eng in ENGINES:                        # (the tc engine cannot certify an all-tie row: exact fallback)
        gi, gs = topkrec.score_topk(t(U), t(V), 30, engine=eng)
        assert np.array_equal(gi.cpu().numpy(), np.tile(np.arange(899, 869, -1, dtype=np.int32), (150, 1)))
        assert (gs.cpu().numpy().view(np.uint32) == 0).all()       # +0.0, never -0.0


def test_score_topk_rated_mask_and_short_lists():
    U, V, b, p, i = _case(120, 60, 32, 30, seed=12, bias=True, rated=45)   # some users have < 30 unrated columns
    _check(U, V, 30, b, p, i)
    full = np.arange(60, dtype=np.int32)                                    # a user who rated everything
    p2 = np.array([0, 60], np.int64)
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    for eng in ENGINES:
        gi, gs = topkrec.score_topk(t(U[:1]), t(V), 30, None, t(p2), t(full), engine=eng)
        assert (gi.cpu().numpy() == -1).all() and np.isneginf(gs.cpu().numpy()).all()


def test_score_topk_split_path_and_col_offset():
    """few users x many items takes the item-split + merge path; shards report global columns"""
    U, V, b, p, i = _case(40, 20000, 128, 30, seed=13, bias=True, rated=64)
    _check(U, V, 30, b, p, i)
    half = 9984
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()  # noqa: E731
    ri, rs = topk_ref.score_topk(U, V, 30, b, p, i)
    for eng in ENGINES:
        parts = [topkrec.score_topk(t(U), t(V[a:z].copy()), 30, t(b[a:z].copy()), t(p), t(i), col_offset=a, engine=eng)
                 for a, z in ((0, half), (half, 20000))]
        mi, ms = topkrec.topk_merge(torch.stack([x[0] for x in parts]), torch.stack([x[1] for x in parts]))
        assert np.array_equal(mi.cpu().numpy(), ri) and np.array_equal(ms.cpu().numpy().view(np.uint32), rs.view(np.uint32))


def test_topk_merge_matches_oracle():
    rng = np.random.default_rng(14)
    G, nu, k = 8, 257, 30
    score = -np.sort(-rng.standard_normal((G, nu, k)).astype(np.float32), axis=2)
    score[:, :, 20:][rng.random((G, nu, 10)) < 0.3] = 0.25            # ties across lists
    score = -np.sort(-score, axis=2)
    idx = np.empty((G, nu, k), np.int32)
    for g in range(G):
        for r in range(nu):
            c = rng.choice(1000, k, replace=False) + 1000 * g
            o = np.lexsort((-c, -score[g, r]))                           # (score desc, col desc) inside a list
            idx[g, r] = c[o]; score[g, r] = score[g, r][o]
    idx[1, :, 25:] = -1; score[1, :, 25:] = -np.inf                      # a short list
    idx[2, 5, :] = -1; score[2, 5, :] = -np.inf                          # an empty list
    mi, ms = topkrec.topk_merge(torch.from_numpy(idx).cuda(), torch.from_numpy(score).cuda())
    ri, rs = topk_ref.topk_merge(idx, score)
    assert np.array_equal(mi.cpu().numpy(), ri) and np.array_equal(ms.cpu().numpy(), rs)


def test_host_entry_equals_device_entry():
    U, V, b, p, i = _case(300, 3000, 128, 30, seed=15, bias=True, rated=30)
    hi, hs = topkrec.score_topk_host(U, V, 30, b, p, i)
    ri, rs = topk_ref.score_topk(U, V, 30, b, p, i)
    assert np.array_equal(hi, ri) and np.array_equal(hs.view(np.uint32), rs.view(np.uint32))
    hi, hs = topkrec.score_topk_host(U, V, 30)
    ri, rs = topk_ref.score_topk(U, V, 30)
    assert np.array_equal(hi, ri)


def test_evaluate_cli_matches_reference_goldens(golden, mini, capsys):
    """our evaluate.py (reference CLI) on the reference-written .dat models == the numbers the
    unmodified reference script printed (tests/golden/evaluate_mini.json)"""
    import importlib.util
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("tkr_evaluate", os.path.join(PKG, "evaluate.py"))
    ev = importlib.util.module_from_spec(spec); spec.loader.exec_module(ev)
    g = json.load(open(os.path.join(golden, "evaluate_mini.json")))
    assert ev.main(["-d", mini, "-m", os.path.join(golden, "mini_model"), "-sl", "im", "om", "all"]) == g["no_bias"]
    assert ev.main(["-d", mini, "-m", os.path.join(golden, "mini_model_bias"), "-sl", "all"]) == g["bias_all"]
    assert ev.main(["-d", mini, "-m", os.path.join(golden, "mini_model_ties"), "-sl", "im", "om", "all"]) == g["ties_oracle_stable"]
    out = capsys.readouterr().out.strip().splitlines()
    assert out[:3] == g["no_bias"]


def test_evaluate_lists_match_golden_lists(golden, mini):
    import utils
    lists = np.load(os.path.join(golden, "evaluate_mini_lists.npz"))
    uids = utils.get_id_dict_from_file(os.path.join(mini, "uid")); vids = utils.get_id_dict_from_file(os.path.join(mini, "vid"))
    browsed, _ = utils.get_history_from_file(os.path.join(mini, "f0tr.txt"))
    for model, key_prefix in (("mini_model", ""), ("mini_model_ties", "ties_")):
        U = utils.get_embed_from_file(os.path.join(golden, model, "final-U.dat"), uids)
        V = utils.get_embed_from_file(os.path.join(golden, model, "final-V.dat"), vids)
        for sc in ("im", "om", "all"):
            teids = utils.get_id_dict_from_file(os.path.join(mini, "f0te.%s.idl" % sc))
            cols = np.array([vids[v] for v in teids])
            p, i = utils.rated_csr(uids, browsed, teids)
            hi, _ = topkrec.score_topk_host(U, V[cols], 30, None, p, i)
            assert np.array_equal(hi, lists[key_prefix + sc]), (model, sc)


def test_eval_hits_matches_oracle_walk(golden, mini):
    """tkr_eval_hits (evaluate.py:84-112 on the device) == the oracle's Python walk over the golden lists"""
    import importlib.util
    import utils
    from conftest import PKG
    from oracle import evaluate_ref
    spec = importlib.util.spec_from_file_location("tkr_evaluate", os.path.join(PKG, "evaluate.py"))
    ev = importlib.util.module_from_spec(spec); spec.loader.exec_module(ev)
    lists = np.load(os.path.join(golden, "evaluate_mini_lists.npz"))
    uids = utils.get_id_dict_from_file(os.path.join(mini, "uid"))
    for key in ("im", "om", "all", "ties_im", "ties_all"):
        sc = key.split("_")[-1]
        teids = utils.get_id_dict_from_file(os.path.join(mini, "f0te.%s.idl" % sc))
        te_file = os.path.join(mini, "f0te.%s.txt" % sc)
        for step, total in ((5, 30), (1, 30), (7, 28), (10, 20)):
            L = np.ascontiguousarray(lists[key][:, :total])
            ref_hits, ref_cnt = evaluate_ref.hits_from_lists(L, uids, teids, te_file, step, total)
            hits, cnt = ev.count_hits(torch.from_numpy(L).cuda(), os.path.join(mini, "uid"), te_file, os.path.join(mini, "f0te.%s.idl" % sc), step, total)
            assert cnt == ref_cnt and np.array_equal(hits, ref_hits), (key, step, total)


def test_eval_hits_random_large():
    """synthetic: 50k lines x top-30, ragged like lists (some empty-list rows, -1 padding), vs numpy"""
    rng = np.random.default_rng(21)
    n_rows, n_cols, total, n_lines, step = 20000, 5000, 30, 50000, 5
    lists = np.stack([rng.choice(n_cols, total, replace=False) for _ in range(n_rows)]).astype(np.int32)
    lists[rng.random(n_rows) < 0.1, 20:] = -1
    rows = rng.integers(0, n_rows, n_lines).astype(np.int32)
    cnt = rng.integers(1, 40, n_lines)
    indptr = np.zeros(n_lines + 1, np.int64); indptr[1:] = np.cumsum(cnt)
    idx = np.concatenate([np.sort(rng.choice(n_cols, c, replace=False)) for c in cnt]).astype(np.int32)
    pos = np.zeros(total, np.int64)
    for l in range(0, n_lines, 1):
        row = lists[rows[l]]
        m = np.isin(row, idx[indptr[l]:indptr[l + 1]]) & (row >= 0)
        pos += m
    ref = np.array([pos[:(q + 1) * step].sum() for q in range(total // step)], np.float64)
    hits, ph = topkrec.eval_hits(torch.from_numpy(lists).cuda(), torch.from_numpy(rows).cuda(), torch.from_numpy(indptr).cuda(),
                                 torch.from_numpy(idx).cuda(), step)
    assert np.array_equal(ph.cpu().numpy(), pos) and np.array_equal(hits, ref)


def test_full_size_property_sharded_equals_whole():
    """C5-like width (1M items, d=128, k=30) on a small user batch: splitting the items into 8 shards and
    merging gives the same bits as the single call; the returned scores are the exact FMA-chain scores;
    lists are sorted by (score desc, col desc)."""
    rng = np.random.default_rng(16)
    nu, ni, d, k = 256, 1 << 20, 128, 30
    U = torch.from_numpy((0.1 * rng.standard_normal((nu, d))).astype(np.float32)).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    V = torch.randn(ni, d, device="cuda", generator=g) * 0.1
    wi, ws_ = topkrec.score_topk(U, V, k, engine="tc")
    ei, es = topkrec.score_topk(U, V, k, engine="exact")
    assert torch.equal(ei, wi) and torch.equal(es, ws_), "tensor-core path and exact engine disagree at full width"
    bounds = np.linspace(0, ni, 9).astype(np.int64)
    parts = [topkrec.score_topk(U, V[a:z], k, col_offset=int(a), engine="tc") for a, z in zip(bounds[:-1], bounds[1:])]
    mi, ms = topkrec.topk_merge(torch.stack([x[0] for x in parts]), torch.stack([x[1] for x in parts]))
    assert torch.equal(mi, wi) and torch.equal(ms, ws_)
    wi_h, ws_h = wi.cpu().numpy(), ws_.cpu().numpy()
    assert (np.diff(ws_h, axis=1) <= 0).all()
    Vh = V[torch.from_numpy(wi_h[:8].ravel().astype(np.int64)).cuda()].cpu().numpy().reshape(8, k, d)
    Uh = U[:8].cpu().numpy()
    for r in range(8):
        _, rs = topk_ref.score_topk(Uh[r:r + 1], Vh[r], k)
        assert np.array_equal(np.sort(rs[0])[::-1].view(np.uint32), ws_h[r].view(np.uint32))
    # nothing outside the list beats the k-th score (fp64 check with a margin >> fp32 dot error)
    S = (U[:8].double() @ V.double().T)
    kth = torch.from_numpy(ws_h[:8, -1]).cuda().double()
    assert int((S > (kth[:, None] + 1e-4)).sum().item()) <= 8 * (k - 1)


def test_tc_adversarial_rows_fall_back_and_stay_exact():
    """near-ties at the k-th place (gap << bf16 error) cannot be certified: those rows must take the exact
    fallback and still come out bit-identical; a wide user batch takes the single-split path."""
    rng = np.random.default_rng(21)
    nu, ni, d, k = 700, 6000, 128, 30
    U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
    V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
    V[1000:1100] = V[999] * (1 + 1e-6 * rng.standard_normal((100, 1))).astype(np.float32)   # 100 near-duplicates
    fb = _check(U, V, k, None, None, None, engines=("tc",))
    assert fb["tc"] > 0
    big = _case(19000, 3000, 64, 30, seed=22, bias=True, rated=20)
    _check(big[0], big[1], 30, big[2], big[3], big[4], engines=("tc",))


def test_tc_large_norm_spread():
    """rows with very different norms: the per-row error bound scales with |u| * max|v|"""
    rng = np.random.default_rng(23)
    U = (rng.standard_normal((400, 96)) * np.exp(rng.uniform(-6, 3, (400, 1)))).astype(np.float32)
    V = (rng.standard_normal((9000, 96)) * np.exp(rng.uniform(-6, 3, (9000, 1)))).astype(np.float32)
    _check(U, V, 30, None, None, None)


def test_tc_fallback_rows_use_item_splits():
    """uncertified rows over a wide item table go through the item-split fallback (+ merge through the row map)
    and must still be bit-identical; > 1024 failing rows also exercises the unsplit tail."""
    rng = np.random.default_rng(24)
    nu, ni, d, k = 1500, 40000, 64, 30
    U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
    V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
    V[5000:5200] = V[4999]                                   # 200 exact duplicates: every row that ranks them high ties at the cut
    U[:1200] = np.abs(U[:1200]) * np.sign(V[4999])           # make the duplicate block score high for 1200 rows
    fb = _check(U, V, k, None, None, None, engines=("tc",))
    assert fb["tc"] >= 1100


def test_tc_items_prepared_reuses_bf16_table():
    """an evaluator scores several user batches against one item table: the second batch reuses the converted items"""
    U, V, b, p, i = _case(600, 5000, 128, 30, seed=25, bias=True, rated=30)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()  # noqa: E731
    need = topkrec.lib().tkr_score_topk_tc_workspace_bytes(300, 5000, 128, 30, 1)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    Vd, bd, idx_d = t(V), t(b), t(i)
    for h, prepared in ((0, False), (1, True)):
        rows = slice(300 * h, 300 * (h + 1))
        ptr = p[300 * h:300 * (h + 1) + 1]
        gi, gs = topkrec.score_topk(t(U[rows].copy()), Vd, 30, bd, t(ptr.copy()), idx_d, engine="tc", ws=ws, items_prepared=prepared)
        ri, rs = topk_ref.score_topk(U[rows], V, 30, b, ptr - ptr[0], i[ptr[0]:ptr[-1]])
        assert np.array_equal(gi.cpu().numpy(), ri) and np.array_equal(gs.cpu().numpy().view(np.uint32), rs.view(np.uint32))


@pytest.mark.parametrize("nu,ni,d,bias,rated,cuts", [(18944, 65536, 128, False, 0, (0, 20000, 41000, 65536)), (19000, 40000, 64, True, 40, (0, 9000, 40000)),
                                                    (18944, 30000, 200, False, 16, (0, 7000, 14000, 22000, 30000)),
                                                    (18944, 100000, 64, True, 24, (0, 25000, 60000, 100000))])
def test_score_topk_sweep_in_segments_equals_whole_sweep(nu, ni, d, bias, rated, cuts):
    """tkr_score_topk_tc_segment: one sweep cut into segments over item shards, the rows' thresholds / candidate buffers carried
    in the state from segment to segment (what travels from GPU to GPU in the ring) -- lists and score bits equal the oracle's
    and the single-call engine's on the whole table.  Tables of >= 65 536 items: the first segment seeds its thresholds on a
    sample of the WHOLE table (BF16 copy in the segment workspace), with and without a bias column"""
    U, V, b, indptr, idx = _case(nu, ni, d, 30, seed=nu + ni + d, bias=bias, rated=rated)
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()  # noqa: E731
    Ud, Vd, bd, rp, ri = t(U), t(V), t(b), t(indptr), t(idx)
    state = torch.zeros(topkrec.lib().tkr_score_topk_tc_state_bytes(nu), dtype=torch.uint8, device="cuda")
    nfb = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = None
    for s in range(len(cuts) - 1):
        lo, hi = cuts[s], cuts[s + 1]
        out = topkrec.score_topk_segment(Ud, Vd[lo:hi].contiguous(), 30, lo, state, s == 0, s == len(cuts) - 2, V_full=Vd,
                                         bias_shard=None if bd is None else bd[lo:hi].contiguous(), bias_full=bd, rated_indptr=rp, rated_idx=ri, n_fallback=nfb)
    gi, gs = out[0].cpu().numpy(), out[1].cpu().numpy()
    wi, wsc = topkrec.score_topk(Ud, Vd, 30, bd, rp, ri, engine="tc")
    assert np.array_equal(gi, wi.cpu().numpy()) and np.array_equal(gs.view(np.uint32), wsc.cpu().numpy().view(np.uint32))
    rows = np.random.default_rng(1).choice(nu, 300, replace=False)
    sub_ptr = None if indptr is None else np.concatenate([[0], np.cumsum(np.diff(indptr)[rows])]).astype(np.int64)
    sub_idx = None if indptr is None else np.concatenate([idx[indptr[r]:indptr[r + 1]] for r in rows] + [np.zeros(0, np.int32)]).astype(np.int32)
    oi, osc = topk_ref.score_topk(U[rows], V, 30, b, sub_ptr, sub_idx)
    assert np.array_equal(gi[rows], oi) and np.array_equal(gs[rows].view(np.uint32), osc.view(np.uint32))
    assert int(nfb.item()) <= nu // 100

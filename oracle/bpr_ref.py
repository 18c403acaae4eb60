"""numpy restatement of the BPR / VBPR mini-batch step (hot path 1).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY UNPINNED: the step
arithmetic of the reference runs inside TensorFlow 1.15 (``requirements.txt:3``),
which is not vendored and cannot be installed here.  Cross-checked (tests/test_oracle.py)
by fp64 finite differences, a hand-computed duplicate case, and -- independently of the
hand-derived gradients below -- by ``oracle/tf_literal.py`` (the reference's graph lines in
torch autograd + a statement-by-statement transcription of TF's RMSProp kernels and
duplicate-index plumbing).  This file restates

* the objective and its per-occurrence gradients   ``single/bpr.py:81-99``
* ``RMSPropOptimizer(lr).minimize(obj)``           ``single/bpr.py:100``
  = gradient -> concat of the two item gathers -> unique + segment-sum
    (``Optimizer._apply_sparse_duplicate_indices``) -> ``SparseApplyRMSProp``
    with decay 0.9, momentum 0, epsilon 1e-10 *inside* the sqrt, ``rms`` slot
    initialised to ones (TF r1.15 ``python/training/rmsprop.py``,
    ``core/kernels/training_ops.cc``)
* the VBPR variant                                  ``single/vbpr.py:50-73``
* the legacy plain-SGD update                       ``old/methods/bpr.py:43-62``

Everything is vectorised numpy in the dtype of the state arrays (fp32 for the
parity runs, fp64 for the shadow / finite-difference checks).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RMS_DECAY = 0.9
RMS_EPS = 1e-10


@dataclass
class BprCfg:
    """Hyper-parameters; defaults are ``single/bpr.py:20``."""
    lambda_u: float = 2.5e-3
    lambda_i: float = 2.5e-3
    lambda_j: float = 2.5e-4
    lambda_b: float = 0.0
    lambda_e: float = 0.0          # VBPR only (``single/vbpr.py:18``)
    lr: float = 1.0e-4
    mode: str = "l2"               # anything else = L1 (``single/bpr.py:96-99``)
    optimizer: str = "rmsprop"     # "sgd" = ``old/methods/bpr.py:57-61``


def new_state(n_users, n_items, k, rng, dtype=np.float32):
    """Initial state per ``single/bpr.py:77-79`` (N(0, 0.01), bias 0, rms 1)."""
    st = {
        "U": (0.01 * rng.standard_normal((n_users, k))).astype(dtype),
        "V": (0.01 * rng.standard_normal((n_items, k))).astype(dtype),
        "b": np.zeros(n_items, dtype),
    }
    st["msU"] = np.ones_like(st["U"])
    st["msV"] = np.ones_like(st["V"])
    st["msb"] = np.ones_like(st["b"])
    return st


def _reg_value(x, lam, l2):
    return 0.5 * lam * np.sum(x * x) if l2 else lam * np.sum(np.abs(x))


def _reg_grad(x, lam, l2):
    return lam * x if l2 else lam * np.sign(x)


def bpr_forward(U, V, b, u, i, j, cfg: BprCfg):
    """x_uij and the batch objective, ``single/bpr.py:81-99``."""
    l2 = cfg.mode == "l2"
    Uu, Vi, Vj = U[u], V[i], V[j]
    bi, bj = b[i], b[j]
    x_ui = np.sum(Uu * Vi, axis=1)
    x_uj = np.sum(Uu * Vj, axis=1)
    x = bi - bj + x_ui - x_uj
    loss = np.sum(np.log1p(np.exp(-x)))
    loss += _reg_value(Uu, cfg.lambda_u, l2) + _reg_value(Vi, cfg.lambda_i, l2) + _reg_value(Vj, cfg.lambda_j, l2)
    loss += _reg_value(bi, cfg.lambda_b, l2) + _reg_value(bj, cfg.lambda_b, l2)
    return x, loss


def bpr_occurrence_grads(U, V, b, u, i, j, cfg: BprCfg):
    """Per-occurrence gradients at the pre-step snapshot."""
    l2 = cfg.mode == "l2"
    dt = U.dtype
    Uu, Vi, Vj = U[u], V[i], V[j]
    bi, bj = b[i], b[j]
    x, loss = bpr_forward(U, V, b, u, i, j, cfg)
    s = (1.0 / (1.0 + np.exp(x))).astype(dt)          # sigma(-x) = e^-x / (1 + e^-x)
    sc = s[:, None]
    gU = -sc * (Vi - Vj) + _reg_grad(Uu, dt.type(cfg.lambda_u), l2)
    gVi = -sc * Uu + _reg_grad(Vi, dt.type(cfg.lambda_i), l2)
    gVj = sc * Uu + _reg_grad(Vj, dt.type(cfg.lambda_j), l2)
    gbi = -s + _reg_grad(bi, dt.type(cfg.lambda_b), l2)
    gbj = s + _reg_grad(bj, dt.type(cfg.lambda_b), l2)
    return loss, s, gU.astype(dt), gVi.astype(dt), gVj.astype(dt), gbi.astype(dt), gbj.astype(dt)


def segment_sum(indices, values):
    """``_deduplicate_indexed_slices``: unique rows + summed duplicates."""
    rows, inv = np.unique(indices, return_inverse=True)
    out = np.zeros((rows.shape[0],) + values.shape[1:], values.dtype)
    np.add.at(out, inv, values)
    return rows, out


def apply_sparse(var, ms, rows, G, cfg: BprCfg):
    """``SparseApplyRMSProp`` on the unique rows (App. A.4) or plain SGD."""
    dt = var.dtype.type
    if cfg.optimizer == "sgd":
        var[rows] = var[rows] - dt(cfg.lr) * G
        return
    m = dt(RMS_DECAY) * ms[rows] + dt(1.0 - RMS_DECAY) * G * G
    ms[rows] = m
    var[rows] = var[rows] - dt(cfg.lr) * G / np.sqrt(m + dt(RMS_EPS))


def apply_dense(var, ms, G, cfg: BprCfg):
    dt = var.dtype.type
    if cfg.optimizer == "sgd":
        var -= dt(cfg.lr) * G
        return
    ms *= dt(RMS_DECAY)
    ms += dt(1.0 - RMS_DECAY) * G * G
    var -= dt(cfg.lr) * G / np.sqrt(ms + dt(RMS_EPS))


def bpr_step(st, u, i, j, cfg: BprCfg):
    """One synchronous mini-batch step, in place on ``st``; returns the loss
    of the batch evaluated *before* the update (what ``sess.run([solver, obj])``
    returns, ``single/bpr.py:141``)."""
    U, V, b = st["U"], st["V"], st["b"]
    u = np.asarray(u, np.int64); i = np.asarray(i, np.int64); j = np.asarray(j, np.int64)
    loss, _s, gU, gVi, gVj, gbi, gbj = bpr_occurrence_grads(U, V, b, u, i, j, cfg)
    rU, GU = segment_sum(u, gU)
    ij = np.concatenate([i, j])
    rV, GV = segment_sum(ij, np.concatenate([gVi, gVj]))
    rB, GB = segment_sum(ij, np.concatenate([gbi, gbj]))
    apply_sparse(U, st["msU"], rU, GU, cfg)
    apply_sparse(V, st["msV"], rV, GV, cfg)
    apply_sparse(b, st["msb"], rB, GB, cfg)
    return float(loss)


def bpr_train(st, u, i, j, batch_size, cfg: BprCfg):
    """Run ``len(u)//batch_size`` consecutive steps over a flat triple stream."""
    n_steps = len(u) // batch_size
    losses = np.zeros(n_steps, np.float64)
    for t in range(n_steps):
        sl = slice(t * batch_size, (t + 1) * batch_size)
        losses[t] = bpr_step(st, u[sl], i[sl], j[sl], cfg)
    return losses


# --------------------------------------------------------------------------
# VBPR (``single/vbpr.py:29-74``)
# --------------------------------------------------------------------------

def new_vbpr_state(n_users, n_items, k, d, rng, dtype=np.float32):
    """``single/vbpr.py:37-48``: N(0,0.01) embeddings, E = 2/(d k), biases 0."""
    h = k // 2
    st = {
        "UR": (0.01 * rng.standard_normal((n_users, h))).astype(dtype),
        "UC": (0.01 * rng.standard_normal((n_users, h))).astype(dtype),
        "IR": (0.01 * rng.standard_normal((n_items, h))).astype(dtype),
        "rb": np.zeros(n_items, dtype),
        "E": np.full((d, h), 2.0 / (d * k), dtype),
        "c": np.zeros(d, dtype),
    }
    for name in list(st):
        st["ms" + name] = np.ones_like(st[name])
    return st


def _vbpr_parts(st, F, u, i, j):
    """r_n = rb[i]-rb[j] + (F[i]-F[j]).c  (the [B,1] terms of ``vbpr.py:61``) and y_n = x_ui - x_uj (the [B] terms)."""
    ur, uc = st["UR"][u], st["UC"][u]
    Fi, Fj = F[i], F[j]
    ice, jce = Fi @ st["E"], Fj @ st["E"]
    y = np.sum(ur * st["IR"][i] + uc * ice, axis=1) - np.sum(ur * st["IR"][j] + uc * jce, axis=1)
    r = st["rb"][i] - st["rb"][j] + (Fi - Fj) @ st["c"]
    return r, y


def _vbpr_reg(st, u, i, j, cfg: BprCfg):
    l2 = cfg.mode == "l2"
    reg = _reg_value(st["E"], cfg.lambda_e, l2)
    reg += _reg_value(st["UR"][u], cfg.lambda_u, l2) + _reg_value(st["UC"][u], cfg.lambda_u, l2)
    reg += _reg_value(st["IR"][i], cfg.lambda_i, l2) + _reg_value(st["IR"][j], cfg.lambda_j, l2)
    reg += _reg_value(st["rb"][i], cfg.lambda_b, l2) + _reg_value(st["rb"][j], cfg.lambda_b, l2) + _reg_value(st["c"], cfg.lambda_b, l2)
    return reg


def vbpr_forward(st, F, u, i, j, cfg: BprCfg, pairwise=False):
    """x and the batch objective of ``single/vbpr.py:59-72``.  ``pairwise=False``: the per-triple x_n = r_n + y_n the
    code evidently means.  ``pairwise=True``: the graph as written -- the bias variables are ``[n,1]`` (``vbpr.py:43,47``)
    so ``vbpr.py:61`` broadcasts to x[a,b] = r_a + y_b and the loss sums over all B*B entries (defect D-14)."""
    r, y = _vbpr_parts(st, F, u, i, j)
    x = r[:, None] + y[None, :] if pairwise else r + y
    return x, np.sum(np.log1p(np.exp(-x))) + _vbpr_reg(st, u, i, j, cfg)


def vbpr_step(st, F, u, i, j, cfg: BprCfg, pairwise=False):
    """One VBPR step: sparse RMSProp on UR/UC/IR/rb, dense on E/c.  With ``pairwise`` the weight of triple n is
    sum_a sigma(-x[a,n]) on everything reached through y (embeddings, E) and sum_b sigma(-x[n,b]) on everything reached
    through r (rb, c); per-triple they are both sigma(-x_n)."""
    l2 = cfg.mode == "l2"
    dt = st["UR"].dtype
    u = np.asarray(u, np.int64); i = np.asarray(i, np.int64); j = np.asarray(j, np.int64)
    x, loss = vbpr_forward(st, F, u, i, j, cfg, pairwise)
    sig = (1.0 / (1.0 + np.exp(x))).astype(dt)
    s, sb = (sig.sum(axis=0).astype(dt), sig.sum(axis=1).astype(dt)) if pairwise else (sig, sig)
    sc = s[:, None]
    ur, uc = st["UR"][u], st["UC"][u]
    iri, irj = st["IR"][i], st["IR"][j]
    bi, bj = st["rb"][i], st["rb"][j]
    dF = (F[i] - F[j]).astype(dt)
    g_ur = -sc * (iri - irj) + _reg_grad(ur, dt.type(cfg.lambda_u), l2)
    g_uc = -sc * (dF @ st["E"]) + _reg_grad(uc, dt.type(cfg.lambda_u), l2)
    g_iri = -sc * ur + _reg_grad(iri, dt.type(cfg.lambda_i), l2)
    g_irj = sc * ur + _reg_grad(irj, dt.type(cfg.lambda_j), l2)
    g_bi = -sb + _reg_grad(bi, dt.type(cfg.lambda_b), l2)
    g_bj = sb + _reg_grad(bj, dt.type(cfg.lambda_b), l2)
    g_E = dF.T @ (-sc * uc) + _reg_grad(st["E"], dt.type(cfg.lambda_e), l2)
    g_c = dF.T @ (-sb) + _reg_grad(st["c"], dt.type(cfg.lambda_b), l2)
    rU, G_ur = segment_sum(u, g_ur.astype(dt))
    _, G_uc = segment_sum(u, g_uc.astype(dt))
    ij = np.concatenate([i, j])
    rV, G_ir = segment_sum(ij, np.concatenate([g_iri, g_irj]).astype(dt))
    _, G_rb = segment_sum(ij, np.concatenate([g_bi, g_bj]).astype(dt))
    apply_sparse(st["UR"], st["msUR"], rU, G_ur, cfg)
    apply_sparse(st["UC"], st["msUC"], rU, G_uc, cfg)
    apply_sparse(st["IR"], st["msIR"], rV, G_ir, cfg)
    apply_sparse(st["rb"], st["msrb"], rV, G_rb, cfg)
    apply_dense(st["E"], st["msE"], g_E.astype(dt), cfg)
    apply_dense(st["c"], st["msc"], g_c.astype(dt), cfg)
    return float(loss)


def vbpr_export(st, F):
    """``single/vbpr.py:124-126``: fold the content part into (fue, fie, fib)."""
    fue = np.concatenate([st["UR"], st["UC"]], axis=1)
    fie = np.concatenate([st["IR"], F @ st["E"]], axis=1)
    fib = (st["rb"] + F @ st["c"]).reshape(-1, 1)
    return fue, fie, fib


# --------------------------------------------------------------------------
# ctypes face of oracle/bpr_ref.c (the same step in C + OpenMP, for full-size runs)
# --------------------------------------------------------------------------

def c_bpr_train(st, u, i, j, batch_size, cfg: BprCfg, n_users=None, n_items=None):
    """``bpr_train`` through ``tkr_ref_bpr_step`` (fp32 state arrays updated in place); returns the per-step losses."""
    import ctypes
    from . import clib

    class _Cfg(ctypes.Structure):
        _fields_ = [("n_users", ctypes.c_int32), ("n_items", ctypes.c_int32), ("d", ctypes.c_int32),
                    ("lu", ctypes.c_float), ("li", ctypes.c_float), ("lj", ctypes.c_float), ("lb", ctypes.c_float),
                    ("lr", ctypes.c_float), ("l1", ctypes.c_int32), ("sgd", ctypes.c_int32)]
    for n in ("U", "V", "b", "msU", "msV", "msb"):
        assert st[n].dtype == np.float32 and st[n].flags.c_contiguous, n
    u, i, j = (np.ascontiguousarray(a, np.int32) for a in (u, i, j))
    nu, d = st["U"].shape
    c = _Cfg(nu, st["V"].shape[0], d, cfg.lambda_u, cfg.lambda_i, cfg.lambda_j, cfg.lambda_b, cfg.lr,
             int(cfg.mode != "l2"), int(cfg.optimizer == "sgd"))
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    n_steps = len(u) // batch_size
    losses, loss = np.zeros(n_steps), ctypes.c_double()
    step = clib.lib().tkr_ref_bpr_step
    for t in range(n_steps):
        o = t * batch_size
        rc = step(ctypes.byref(c), fp(st["U"]), fp(st["V"]), fp(st["b"]), fp(st["msU"]), fp(st["msV"]), fp(st["msb"]),
                  fp(u[o:o + batch_size]), fp(i[o:o + batch_size]), fp(j[o:o + batch_size]), ctypes.c_int64(batch_size), ctypes.byref(loss))
        assert rc == 0
        losses[t] = loss.value
    return losses

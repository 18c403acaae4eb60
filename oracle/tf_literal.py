"""A SECOND, independent statement of hot path 1, used only to cross-check ``oracle/bpr_ref.py``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ``bpr_ref`` derives the gradients by hand and
restates RMSProp as one formula.  This module does neither:

* the objective is the reference's graph, transcribed line by line into torch ops with the same broadcasting rules
  (``single/bpr.py:81-99``, ``single/vbpr.py:50-72``) and differentiated by ``torch.autograd`` -- the gradient of a
  gather is a scatter-add, i.e. the summed ``IndexedSlices`` TensorFlow builds for ``embedding_lookup``;
* the update is a literal, statement-by-statement transcription of the TensorFlow r1.15 optimiser plumbing
  (``python/training/optimizer.py``: ``_deduplicate_indexed_slices`` / ``_apply_sparse_duplicate_indices``;
  ``python/training/rmsprop.py``: slots ``rms`` = ones, ``momentum`` = zeros, decay 0.9, momentum 0.0, epsilon 1e-10)
  and of its CPU kernels (``core/kernels/training_ops.cc``: ``SparseApplyRMSPropOp::Compute`` and
  ``functor::ApplyRMSProp<CPUDevice, T>``), *including* the momentum slot the closed form drops.

TensorFlow itself is still absent (``requirements.txt:3`` pins 1.15, not installable offline), so path 1 stays
"parity unpinned"; what this buys is that two derivations that share no code agree to rounding.

Reference quirk reproduced here (D-14, found while transcribing): in ``single/vbpr.py`` the bias variables are
``[n_items, 1]`` / ``[d, 1]``, so ``irbb - jrbb`` and ``tf.matmul(ic - jc, icb)`` are ``[B, 1]`` while ``x_ui - x_uj``
is ``[B]``; their sum (``vbpr.py:61``) broadcasts to a ``[B, B]`` matrix ``x[a, b] = r_a + y_b`` and the loss sums
``log(1 + exp(-x))`` over all B*B entries.  ``vbpr_objective(..., pairwise=True)`` is that literal graph,
``pairwise=False`` the per-triple objective the code evidently meant (and BPR, whose bias is ``[n_items]``, has).
"""
from __future__ import annotations

import numpy as np
import torch

RHO, MOMENTUM, EPSILON = 0.9, 0.0, 1e-10      # tf.train.RMSPropOptimizer defaults (rmsprop.py __init__)


# ----------------------------------------------------------------------------- the graphs, line by line
def bpr_objective(ue, ie, ib, u, i, j, lu, li, lj, lb, mode="l2"):
    """``single/bpr.py:81-99``; ue [n_users,k], ie [n_items,k], ib [n_items] (torch, requires_grad)."""
    ueb, ieb, jeb = ue[u], ie[i], ie[j]                     # tf.nn.embedding_lookup, :81-83
    ibb, jbb = ib[i], ib[j]                                 # :84-85
    x_ui = torch.sum(ueb * ieb, 1)                          # :87
    x_uj = torch.sum(ueb * jeb, 1)                          # :88
    x_uij = ibb - jbb + x_ui - x_uj                         # :89
    if mode == "l2":                                        # :92-95
        return (torch.sum(torch.log(1 + torch.exp(-x_uij)))
                + 0.5 * torch.sum(ueb ** 2 * lu + ieb ** 2 * li + jeb ** 2 * lj)
                + 0.5 * torch.sum(ibb ** 2 + jbb ** 2) * lb)
    return (torch.sum(torch.log(1 + torch.exp(-x_uij)))     # :96-99
            + torch.sum(torch.abs(ueb) * lu + torch.abs(ieb) * li + torch.abs(jeb) * lj)
            + torch.sum(torch.abs(ibb) + torch.abs(jbb)) * lb)


def vbpr_objective(ure, uce, ire, irb, cem, icb, feat, u, i, j, lu, li, lj, lb, le, mode="l2", pairwise=True):
    """``single/vbpr.py:50-72``; irb [n_items,1], cem [d,k/2], icb [d,1], feat [n_items,d] (constant)."""
    ic, jc = feat[i], feat[j]                               # the feed of :114
    ureb, uceb = ure[u], uce[u]                             # :50-51
    ireb, jreb = ire[i], ire[j]                             # :52-53
    irbb, jrbb = irb[i], irb[j]                             # :54-55   -> [B, 1]
    iceb, jceb = ic @ cem, jc @ cem                         # :56-57
    x_ui = torch.sum(ureb * ireb + uceb * iceb, 1)          # :59      -> [B]
    x_uj = torch.sum(ureb * jreb + uceb * jceb, 1)          # :60
    if pairwise:
        x_uij = irbb - jrbb + x_ui - x_uj + (ic - jc) @ icb   # :61 verbatim: [B,1] + [B] + [B,1] -> [B, B]
    else:
        x_uij = (irbb - jrbb)[:, 0] + x_ui - x_uj + ((ic - jc) @ icb)[:, 0]
    if mode == "l2":                                        # :63-67
        return (torch.sum(torch.log(1 + torch.exp(-x_uij)))
                + 0.5 * torch.sum(cem ** 2) * le
                + 0.5 * torch.sum((ureb ** 2 + uceb ** 2) * lu + ireb ** 2 * li + jreb ** 2 * lj)
                + 0.5 * (torch.sum(irbb ** 2 + jrbb ** 2) + torch.sum(icb ** 2)) * lb)
    return (torch.sum(torch.log(1 + torch.exp(-x_uij)))     # :68-72
            + torch.sum(torch.abs(cem)) * le
            + torch.sum((torch.abs(ureb) + torch.abs(uceb)) * lu + torch.abs(ireb) * li + torch.abs(jreb) * lj)
            + (torch.sum(torch.abs(irbb) + torch.abs(jrbb)) + torch.sum(torch.abs(icb))) * lb)


# ----------------------------------------------------------------------------- the optimiser, statement by statement
def deduplicate_indexed_slices(values, indices):
    """optimizer.py ``_deduplicate_indexed_slices``: ``unique_indices, new_index_positions = unique(indices)``;
    ``summed_values = unsorted_segment_sum(values, new_index_positions, shape(unique_indices)[0])``.
    (tf.unique keeps first-appearance order; the order of the unique rows does not matter to the kernel below.)"""
    seen, unique_indices, pos = {}, [], np.empty(len(indices), np.int64)
    for n, ix in enumerate(indices.tolist()):
        if ix not in seen:
            seen[ix] = len(unique_indices)
            unique_indices.append(ix)
        pos[n] = seen[ix]
    summed = np.zeros((len(unique_indices),) + values.shape[1:], values.dtype)
    for n in range(len(indices)):                   # unsorted_segment_sum, one occurrence at a time
        summed[pos[n]] += values[n]
    return summed, np.asarray(unique_indices, np.int64)


def sparse_apply_rms_prop(var, ms, mom, lr, rho, momentum, epsilon, grad, indices):
    """training_ops.cc ``SparseApplyRMSPropOp<T, Tindex>::Compute`` (the loop over the N unique indices):
        ms_  = ms_ * rho + grad_.square() * (1 - rho);
        mom_ = mom_ * momentum + (ms_ + epsilon).rsqrt() * lr * grad_;
        v   -= mom_;
    in the dtype of ``var``."""
    T = var.dtype.type
    lr, rho, momentum, epsilon = T(lr), T(rho), T(momentum), T(epsilon)
    for n, index in enumerate(indices.tolist()):
        g = grad[n]
        ms[index] = ms[index] * rho + np.square(g) * (T(1) - rho)
        mom[index] = mom[index] * momentum + (T(1) / np.sqrt(ms[index] + epsilon)) * lr * g
        var[index] = var[index] - mom[index]


def apply_rms_prop(var, ms, mom, lr, rho, momentum, epsilon, grad):
    """training_ops.cc ``functor::ApplyRMSProp<CPUDevice, T>``:
        ms  += (grad.square() - ms) * (1 - rho);
        mom  = mom * momentum + (grad * lr) / (ms + epsilon).sqrt();
        var -= mom;"""
    T = var.dtype.type
    lr, rho, momentum, epsilon = T(lr), T(rho), T(momentum), T(epsilon)
    ms += (np.square(grad) - ms) * (T(1) - rho)
    mom[...] = mom * momentum + (grad * lr) / np.sqrt(ms + epsilon)
    var -= mom


def new_slots(state):
    """rmsprop.py ``_create_slots``: "rms" = ones, "momentum" = zeros, per variable."""
    return {n: (np.ones_like(v), np.zeros_like(v)) for n, v in state.items()}


def _minimize(state, slots, objective, sparse, lr):
    """``RMSPropOptimizer(lr).minimize(obj)``: gradients -> per variable ``_apply_sparse_duplicate_indices`` (variables
    reached through ``embedding_lookup``: ``sparse[name]`` = every index fed to a lookup of that variable, concatenated
    in graph order) or ``_apply_dense``.  Returns the objective evaluated before the update."""
    t = {n: torch.tensor(v, requires_grad=True) for n, v in state.items()}
    obj = objective(t)
    obj.backward()
    for name, var in state.items():
        g = t[name].grad.numpy()
        rms, mom = slots[name]
        if name in sparse:
            # the IndexedSlices TF hands to the optimiser: one slice per looked-up index.  autograd has already summed
            # them per row, so hand every row's sum to its first occurrence and zeros to the repeats -- after
            # _deduplicate_indexed_slices that is the same summed_values / unique_indices pair.
            idx = np.asarray(sparse[name], np.int64)
            vals = np.zeros((len(idx),) + var.shape[1:], var.dtype)
            first = {}
            for n, ix in enumerate(idx.tolist()):
                if ix not in first:
                    first[ix] = n
                    vals[n] = g[ix]
            summed, uniq = deduplicate_indexed_slices(vals, idx)
            sparse_apply_rms_prop(var, rms, mom, lr, RHO, MOMENTUM, EPSILON, summed, uniq)
        else:
            apply_rms_prop(var, rms, mom, lr, RHO, MOMENTUM, EPSILON, g.astype(var.dtype))
    return float(obj.detach())


def bpr_minimize_step(state, slots, u, i, j, lu, li, lj, lb, lr, mode="l2"):
    """One ``sess.run([solver, obj])`` of ``single/bpr.py:141``.  state: numpy ``ue, ie, ib`` (updated in place)."""
    u, i, j = (torch.as_tensor(np.asarray(a, np.int64)) for a in (u, i, j))
    ij = np.concatenate([i.numpy(), j.numpy()])
    return _minimize(state, slots, lambda t: bpr_objective(t["ue"], t["ie"], t["ib"], u, i, j, lu, li, lj, lb, mode),
                     {"ue": u.numpy(), "ie": ij, "ib": ij}, lr)


def vbpr_minimize_step(state, slots, feat, u, i, j, lu, li, lj, lb, le, lr, mode="l2", pairwise=True):
    """One ``sess.run`` of ``single/vbpr.py:114``.  state: numpy ``ure, uce, ire, irb[n,1], cem, icb[d,1]``."""
    u, i, j = (torch.as_tensor(np.asarray(a, np.int64)) for a in (u, i, j))
    ij = np.concatenate([i.numpy(), j.numpy()])
    f = torch.as_tensor(feat)
    return _minimize(state, slots,
                     lambda t: vbpr_objective(t["ure"], t["uce"], t["ire"], t["irb"], t["cem"], t["icb"], f, u, i, j,
                                              lu, li, lj, lb, le, mode, pairwise),
                     {"ure": u.numpy(), "uce": u.numpy(), "ire": ij, "irb": ij}, lr)

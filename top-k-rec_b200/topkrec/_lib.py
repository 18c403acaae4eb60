from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtopkrec.so")


class TkrError(RuntimeError):
    pass


class tkr_bpr_cfg(C.Structure):
    _fields_ = [("n_users", C.c_int32), ("n_items", C.c_int32), ("d", C.c_int32),
                ("lambda_u", C.c_float), ("lambda_i", C.c_float), ("lambda_j", C.c_float), ("lambda_b", C.c_float),
                ("lr", C.c_float), ("rms_decay", C.c_float), ("rms_eps", C.c_float),
                ("l1", C.c_int32), ("optimizer", C.c_int32)]


class tkr_vbpr_cfg(C.Structure):
    _fields_ = [("base", tkr_bpr_cfg), ("d_feat", C.c_int32), ("lambda_e", C.c_float), ("pairwise", C.c_int32)]


class tkr_sampler(C.Structure):
    _fields_ = [("tr_users", C.c_void_p), ("n_tr_users", C.c_int32), ("pos_indptr", C.c_void_p),
                ("pos_idx", C.c_void_p), ("n_items", C.c_int32), ("seed", C.c_uint64)]


class tkr_als_cfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("ridge", C.c_float), ("lreg", C.c_float),
                ("solve_empty", C.c_int32), ("item_loss", C.c_int32)]


class tkr_als_plan(C.Structure):
    _fields_ = [("n_segs", C.c_int64), ("seg_row", C.c_void_p), ("seg_off", C.c_void_p), ("seg_len", C.c_void_p),
                ("seg_slot", C.c_void_p), ("n_multi", C.c_int64), ("multi_row", C.c_void_p), ("multi_slot0", C.c_void_p),
                ("multi_nslots", C.c_void_p), ("multi_total", C.c_void_p), ("n_slots", C.c_int64)]


_lib = None


def lib():
    """Load the library (once).  No fallback: a missing .so is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise TkrError("libtopkrec.so is not built (%s); run `python top-k-rec_b200/build.py` "
                       "or `python -c 'import __graft_entry__ as g; g.build()'`" % _SO)
    L = C.CDLL(_SO)
    vp, i64, i32, u64, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_size_t
    cfgp, smpp = C.POINTER(tkr_bpr_cfg), C.POINTER(tkr_sampler)
    L.tkr_version.restype = C.c_int
    L.tkr_last_error.restype = C.c_char_p
    L.tkr_launch_count.restype = i64
    L.tkr_reset_launch_count.restype = None
    L.tkr_bpr_workspace_bytes.restype = sz; L.tkr_bpr_workspace_bytes.argtypes = [cfgp, i64]
    L.tkr_bpr_workspace_init.argtypes = [cfgp, i64, vp, sz, vp]
    L.tkr_bpr_workspace_layout.argtypes = [cfgp, i64, C.POINTER(C.c_int64)]
    L.tkr_bpr_workspace_set_hot_items.argtypes = [cfgp, i64, vp, sz, vp, i32, vp]
    L.tkr_bpr_grad.argtypes = [cfgp] + [vp] * 3 + [vp] * 3 + [i64, smpp, u64, vp, vp, sz, i32, vp]
    L.tkr_bpr_apply.argtypes = [cfgp] + [vp] * 6 + [i64, vp, sz, i32, vp]
    L.tkr_bpr_step.argtypes = [cfgp] + [vp] * 6 + [vp] * 3 + [i64, i64, smpp, u64, vp, vp, sz, vp]
    L.tkr_bpr_hogwild.argtypes = [cfgp] + [vp] * 3 + [vp] * 3 + [i64, i64, smpp, u64, vp, vp]
    L.tkr_bpr_step_host.argtypes = [cfgp] + [vp] * 6 + [vp] * 3 + [i64, i64, vp, vp, sz, vp, sz, vp]
    L.tkr_bpr_sample.argtypes = [smpp, u64, i64, vp, vp, vp, vp]
    L.tkr_bpr_dp_layout.argtypes = [cfgp, C.POINTER(C.c_int64)]
    L.tkr_bpr_dp_step.argtypes = [cfgp] + [vp] * 4 + [vp] * 3 + [i64, smpp, u64, vp, vp, sz, vp, u64, vp]
    L.tkr_bpr_dp_status.argtypes = [cfgp, vp, vp]
    vcfgp = C.POINTER(tkr_vbpr_cfg)
    L.tkr_vbpr_workspace_bytes.restype = sz; L.tkr_vbpr_workspace_bytes.argtypes = [vcfgp, i64]
    L.tkr_vbpr_workspace_init.argtypes = [vcfgp, i64, vp, sz, vp]
    L.tkr_vbpr_project.argtypes = [vcfgp] + [vp] * 6 + [vp]
    L.tkr_vbpr_step.argtypes = [vcfgp] + [vp] * 12 + [vp] * 3 + [i64, i64, smpp, u64, vp, vp, sz, vp]
    L.tkr_vbpr_grad.argtypes = [vcfgp] + [vp] * 7 + [vp] * 3 + [i64, smpp, u64, vp, vp, sz, i32, vp]
    L.tkr_vbpr_apply.argtypes = [vcfgp] + [vp] * 10 + [i64, vp, vp, sz, i32, vp]
    L.tkr_vbpr_workspace_layout.argtypes = [vcfgp, i64, C.POINTER(C.c_int64)]
    L.tkr_score_topk_workspace_bytes.restype = sz; L.tkr_score_topk_workspace_bytes.argtypes = [i64, i64, i32, i32]
    L.tkr_score_topk.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, i32, i64, vp, vp, vp, sz, vp]
    L.tkr_score_topk_tc_workspace_bytes.restype = sz; L.tkr_score_topk_tc_workspace_bytes.argtypes = [i64, i64, i32, i32, i32]
    L.tkr_score_topk_tc.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, i32, i64, vp, vp, vp, sz, vp, i32, vp]
    L.tkr_score_topk_host_device_bytes.restype = sz
    L.tkr_score_topk_host_device_bytes.argtypes = [i64, i64, i32, i32, i64]
    L.tkr_score_topk_host.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, i32, vp, vp, vp, sz, vp]
    L.tkr_topk_merge.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
    L.tkr_score_topk_tc_state_bytes.restype = sz; L.tkr_score_topk_tc_state_bytes.argtypes = [i64]
    L.tkr_score_topk_tc_segment_workspace_bytes.restype = sz; L.tkr_score_topk_tc_segment_workspace_bytes.argtypes = [i64, i64, i64, i32, i32, i32]
    L.tkr_score_topk_tc_segment.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, i32, i64, vp, i32, i32, vp, i64, vp, vp, vp, vp, sz, vp, i32, vp]
    L.tkr_peer_flag_bytes.restype = sz; L.tkr_peer_flag_bytes.argtypes = []
    L.tkr_peer_signal_to.argtypes = [vp, sz, i32, i32, u64, vp]
    L.tkr_peer_wait_from.argtypes = [vp, sz, i32, i32, u64, vp]
    L.tkr_peer_status.argtypes = [vp, sz, vp]
    L.tkr_topk_exchange_bytes.restype = sz; L.tkr_topk_exchange_bytes.argtypes = [i64, i32, i32]
    L.tkr_topk_exchange_push.argtypes = [vp, vp, i64, i64, i32, vp, u64, vp]
    L.tkr_topk_exchange_merge.argtypes = [i64, i64, i32, vp, u64, vp, vp, vp]
    L.tkr_topk_exchange_status.argtypes = [i64, i32, vp, vp]
    L.tkr_eval_hits.argtypes = [vp, i32, vp, vp, vp, i64, vp, vp]
    L.tkr_dat_shape.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.tkr_dat_read.argtypes = [C.c_char_p, vp, i64, i64]
    L.tkr_dat_write.argtypes = [C.c_char_p, vp, i64, i64]
    L.tkr_dat_write_f64.argtypes = [C.c_char_p, vp, i64, i64]; L.tkr_dat_write_f64.restype = C.c_int
    L.tkr_ratings_parse.argtypes = [C.c_char_p] * 3 + [C.POINTER(C.c_int64)] * 2 + [vp] * 4
    L.tkr_als_partial_bytes.restype = sz; L.tkr_als_partial_bytes.argtypes = [i32, i64]
    L.tkr_als_gram_workspace_bytes.restype = sz; L.tkr_als_gram_workspace_bytes.argtypes = [i32]
    L.tkr_als_gram.argtypes = [vp, i32, vp, i64, C.c_float, C.c_float, vp, vp, sz, vp]
    L.tkr_als_solve_rows.argtypes = [C.POINTER(tkr_als_cfg), C.POINTER(tkr_als_plan), vp, vp, vp, vp, vp, vp, vp, sz, vp]
    for name in ("tkr_score_topk_tc_segment", "tkr_peer_signal_to", "tkr_peer_wait_from", "tkr_peer_status", "tkr_bpr_hogwild", "tkr_vbpr_grad", "tkr_vbpr_apply", "tkr_vbpr_workspace_layout", "tkr_topk_exchange_push", "tkr_topk_exchange_merge", "tkr_topk_exchange_status", "tkr_bpr_dp_layout", "tkr_bpr_dp_step", "tkr_bpr_dp_status", "tkr_als_gram", "tkr_als_solve_rows", "tkr_bpr_workspace_init", "tkr_bpr_workspace_layout", "tkr_bpr_workspace_set_hot_items", "tkr_bpr_grad", "tkr_bpr_apply", "tkr_bpr_step", "tkr_bpr_step_host", "tkr_bpr_sample", "tkr_vbpr_workspace_init", "tkr_vbpr_project", "tkr_vbpr_step", "tkr_score_topk",
                 "tkr_score_topk_tc", "tkr_score_topk_host", "tkr_topk_merge", "tkr_eval_hits", "tkr_dat_shape", "tkr_dat_read", "tkr_dat_write",
                 "tkr_ratings_parse"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise TkrError("libtopkrec error %d: %s" % (rc, lib().tkr_last_error().decode()))


def version():
    return lib().tkr_version()


def launch_count():
    return int(lib().tkr_launch_count())


def reset_launch_count():
    lib().tkr_reset_launch_count()


def _dev(t, dtype, name):
    """Device pointer of a contiguous CUDA tensor of the expected dtype."""
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TkrError("%s must be a CUDA tensor (no CPU path exists)" % name)
    if t.dtype != dtype or not t.is_contiguous():
        raise TkrError("%s must be contiguous %s, got %s contiguous=%s" % (name, dtype, t.dtype, t.is_contiguous()))
    return t.data_ptr()


def _stream():
    if not torch.cuda.is_available():
        raise TkrError("no CUDA device: libtopkrec has no CPU fallback")
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    if not torch.cuda.is_available():
        raise TkrError("no CUDA device: libtopkrec has no CPU fallback")
    for t in tensors:
        if isinstance(t, torch.Tensor) and not t.is_cuda:
            raise TkrError("got a CPU tensor: libtopkrec has no CPU fallback (move it to the device)")


class BprCfg:
    """Host mirror of tkr_bpr_cfg; defaults are the reference's (single/bpr.py:20)."""

    def __init__(self, n_users, n_items, d, lambda_u=2.5e-3, lambda_i=2.5e-3, lambda_j=2.5e-4, lambda_b=0.0,
                 lr=1.0e-4, mode="l2", optimizer="rmsprop", rms_decay=0.9, rms_eps=1e-10):
        if optimizer not in ("rmsprop", "sgd"):
            raise ValueError("optimizer must be 'rmsprop' or 'sgd'")
        self.c = tkr_bpr_cfg(int(n_users), int(n_items), int(d), lambda_u, lambda_i, lambda_j, lambda_b, lr,
                             rms_decay, rms_eps, 0 if mode == "l2" else 1, 0 if optimizer == "rmsprop" else 1)

    @property
    def ptr(self):
        return C.byref(self.c)


class VbprCfg:
    """Host mirror of tkr_vbpr_cfg (single/vbpr.py:18 defaults); k must be even."""

    def __init__(self, n_users, n_items, k, d_feat, lambda_u=2.5e-3, lambda_i=2.5e-3, lambda_j=2.5e-4, lambda_b=0.0, lambda_e=0.0,
                 lr=1.0e-4, mode="l2", optimizer="rmsprop", pairwise=False):
        """pairwise=True: the graph exactly as the reference writes it (vbpr.py:61 broadcasts x to [B, B]; batch <= 4096)"""
        base = BprCfg(n_users, n_items, k, lambda_u, lambda_i, lambda_j, lambda_b, lr, mode, optimizer).c
        self.c = tkr_vbpr_cfg(base, int(d_feat), float(lambda_e), int(bool(pairwise)))

    @property
    def ptr(self):
        return C.byref(self.c)


class Sampler:
    """Device-resident sampler tables (CSR of positives, ascending within a user)."""

    def __init__(self, tr_users, pos_indptr, pos_idx, n_items, seed, device="cuda"):
        self.tr_users = torch.as_tensor(np.ascontiguousarray(tr_users, np.int32)).to(device)
        self.pos_indptr = torch.as_tensor(np.ascontiguousarray(pos_indptr, np.int64)).to(device)
        self.pos_idx = torch.as_tensor(np.ascontiguousarray(pos_idx, np.int32)).to(device)
        self.c = tkr_sampler(self.tr_users.data_ptr(), int(self.tr_users.numel()), self.pos_indptr.data_ptr(),
                             self.pos_idx.data_ptr(), int(n_items), int(seed))

    @property
    def ptr(self):
        return C.byref(self.c)


def bpr_workspace(cfg: BprCfg, batch, device="cuda"):
    """Allocate and zero the per-step scratch (a uint8 CUDA tensor)."""
    _need_cuda()
    n = lib().tkr_bpr_workspace_bytes(cfg.ptr, int(batch))
    ws = torch.empty(n, dtype=torch.uint8, device=device)
    with torch.cuda.device(ws.device):
        _check(lib().tkr_bpr_workspace_init(cfg.ptr, int(batch), ws.data_ptr(), n, _stream()))
    return ws


WS_FIELDS = ("GU", "cntU", "listU", "n_touched", "GV", "Gb", "tchV", "cntV", "listV", "hotV", "stage", "total")
MAX_HOT = 32


def bpr_set_hot_items(cfg: BprCfg, batch, ws, item_ids):
    """Name up to MAX_HOT popular item rows whose gradients the step sums per thread block in shared memory
    (tkr_bpr_workspace_set_hot_items).  item_ids: host ints, distinct; [] clears the set."""
    ids = np.ascontiguousarray(np.asarray(item_ids, np.int32)[:MAX_HOT])
    with torch.cuda.device(ws.device):
        _check(lib().tkr_bpr_workspace_set_hot_items(cfg.ptr, int(batch), ws.data_ptr(), ws.numel(),
                                                     ids.ctypes.data if ids.size else None, int(ids.size), _stream()))


def popular_items(pos_idx, n_items, n=MAX_HOT):
    """The n items with the most positives (ties: smaller id), from the sampler's positives CSR column array."""
    pos_idx = pos_idx.cpu().numpy() if isinstance(pos_idx, torch.Tensor) else np.asarray(pos_idx)
    cnt = np.bincount(pos_idx, minlength=n_items)
    order = np.lexsort((np.arange(n_items), -cnt))[:n]
    return order[cnt[order] > 0].astype(np.int32)


def bpr_workspace_layout(cfg: BprCfg, batch):
    """Byte offsets of the workspace regions (include/topkrec.h, TKR_WS_*)."""
    off = (C.c_int64 * len(WS_FIELDS))()
    _check(lib().tkr_bpr_workspace_layout(cfg.ptr, int(batch), off))
    return dict(zip(WS_FIELDS, (int(x) for x in off)))


def bpr_item_grad_view(cfg: BprCfg, batch, ws):
    """fp32 view of the contiguous [GV | Gb | tchV] region: what data-parallel training all-reduces."""
    lay = bpr_workspace_layout(cfg, batch)
    n = cfg.c.n_items * cfg.c.d + 2 * cfg.c.n_items
    return ws[lay["GV"]:lay["GV"] + 4 * n].view(torch.float32)


def bpr_grad(cfg: BprCfg, U, V, b, u, i, j, batch, ws, loss=None, sampler=None, first_draw=0, data_parallel=False):
    f32, i32 = torch.float32, torch.int32
    _need_cuda(U, V, b, ws)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_grad(cfg.ptr, _dev(U, f32, "U"), _dev(V, f32, "V"), _dev(b, f32, "b"),
                                  _dev(u, i32, "u"), _dev(i, i32, "i"), _dev(j, i32, "j"), int(batch),
                                  sampler.ptr if sampler is not None else None, int(first_draw), _dev(loss, f32, "loss"),
                                  ws.data_ptr(), ws.numel(), int(bool(data_parallel)), _stream()))


def bpr_apply(cfg: BprCfg, U, V, b, msU, msV, msb, batch, ws, data_parallel=False):
    f32 = torch.float32
    _need_cuda(U, V, b, ws)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_apply(cfg.ptr, _dev(U, f32, "U"), _dev(V, f32, "V"), _dev(b, f32, "b"),
                                   _dev(msU, f32, "msU"), _dev(msV, f32, "msV"), _dev(msb, f32, "msb"), int(batch),
                                   ws.data_ptr(), ws.numel(), int(bool(data_parallel)), _stream()))


DP_FIELDS = ("V", "b", "G0", "G1", "flags", "total")


def bpr_dp_layout(cfg: BprCfg):
    """Byte offsets inside the data-parallel exchange buffer (include/topkrec.h, TKR_DP_*)."""
    off = (C.c_int64 * len(DP_FIELDS))()
    _check(lib().tkr_bpr_dp_layout(cfg.ptr, off))
    return dict(zip(DP_FIELDS, (int(x) for x in off)))


def bpr_dp_step(cfg: BprCfg, U, msU, msV, msb, u, i, j, batch, ws, peers, epoch, loss=None, sampler=None, first_draw=0):
    """tkr_bpr_dp_step: gradient kernel + the fused exchange/update kernel over peer memory.  V and b live in the
    exchange buffer (``peers``: a topkrec.peer.PeerBuffer)."""
    f32, i32 = torch.float32, torch.int32
    _need_cuda(U, ws)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_dp_step(cfg.ptr, _dev(U, f32, "U"), _dev(msU, f32, "msU"), _dev(msV, f32, "msV"), _dev(msb, f32, "msb"),
                                     _dev(u, i32, "u"), _dev(i, i32, "i"), _dev(j, i32, "j"), int(batch),
                                     sampler.ptr if sampler is not None else None, int(first_draw), _dev(loss, f32, "loss"),
                                     ws.data_ptr(), ws.numel(), peers.peers_ptr, int(epoch), _stream()))


def bpr_dp_status(cfg: BprCfg, peers):
    """synchronises the stream; raises if a device-side cross-GPU barrier timed out"""
    with torch.cuda.device(peers.device):
        _check(lib().tkr_bpr_dp_status(cfg.ptr, peers.peers_ptr, _stream()))


def bpr_step(cfg: BprCfg, U, V, b, msU, msV, msb, u, i, j, batch, n_steps, ws, loss=None, sampler=None, first_draw=0):
    f32, i32 = torch.float32, torch.int32
    _need_cuda(U, V, b, ws)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_step(cfg.ptr, _dev(U, f32, "U"), _dev(V, f32, "V"), _dev(b, f32, "b"),
                                  _dev(msU, f32, "msU"), _dev(msV, f32, "msV"), _dev(msb, f32, "msb"),
                                  _dev(u, i32, "u"), _dev(i, i32, "i"), _dev(j, i32, "j"), int(batch), int(n_steps),
                                  sampler.ptr if sampler is not None else None, int(first_draw),
                                  _dev(loss, f32, "loss"), ws.data_ptr(), ws.numel(), _stream()))


def bpr_hogwild(cfg: BprCfg, U, V, b, u, i, j, batch, n_steps, loss=None, sampler=None, first_draw=0):
    """tkr_bpr_hogwild: barrier-free plain-SGD steps, one kernel per step (no workspace, no slots; not bit-reproducible)"""
    f32, i32 = torch.float32, torch.int32
    _need_cuda(U, V, b)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_hogwild(cfg.ptr, _dev(U, f32, "U"), _dev(V, f32, "V"), _dev(b, f32, "b"), _dev(u, i32, "u"), _dev(i, i32, "i"),
                                     _dev(j, i32, "j"), int(batch), int(n_steps), sampler.ptr if sampler is not None else None, int(first_draw),
                                     _dev(loss, f32, "loss"), _stream()))


def bpr_step_host(cfg: BprCfg, U, V, b, msU, msV, msb, u_host, i_host, j_host, batch, n_steps, loss_host, staging, ws):
    """u/i/j/loss are HOST tensors (pinned for async copies); the sess.run seam."""
    f32 = torch.float32
    _need_cuda(U, V, b, ws, staging)
    for t, nm in ((u_host, "u_host"), (i_host, "i_host"), (j_host, "j_host")):
        if t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
            raise TkrError("%s must be a contiguous int32 host tensor" % nm)
    with torch.cuda.device(U.device):
        _check(lib().tkr_bpr_step_host(cfg.ptr, _dev(U, f32, "U"), _dev(V, f32, "V"), _dev(b, f32, "b"),
                                       _dev(msU, f32, "msU"), _dev(msV, f32, "msV"), _dev(msb, f32, "msb"),
                                       u_host.data_ptr(), i_host.data_ptr(), j_host.data_ptr(), int(batch), int(n_steps),
                                       loss_host.data_ptr() if loss_host is not None else None,
                                       staging.data_ptr(), staging.numel(), ws.data_ptr(), ws.numel(), _stream()))


VBPR_STATE = ("U", "V", "rb", "bsum", "E", "c")
VBPR_SLOTS = ("msU", "msV", "msrb", "msE", "msc")


def vbpr_workspace(cfg: VbprCfg, batch, device="cuda"):
    _need_cuda()
    n = lib().tkr_vbpr_workspace_bytes(cfg.ptr, int(batch))
    ws = torch.empty(n, dtype=torch.uint8, device=device)
    with torch.cuda.device(ws.device):
        _check(lib().tkr_vbpr_workspace_init(cfg.ptr, int(batch), ws.data_ptr(), n, _stream()))
    return ws


def vbpr_set_hot_items(cfg: VbprCfg, batch, ws, item_ids):
    """bpr_set_hot_items for a VBPR workspace (it starts with the BPR step workspace): the named popular item rows get
    their gradients -- rating part, content part W and the wq sums -- summed per thread block in shared memory."""
    ids = np.ascontiguousarray(np.asarray(item_ids, np.int32)[:MAX_HOT])
    with torch.cuda.device(ws.device):
        _check(lib().tkr_bpr_workspace_set_hot_items(C.byref(cfg.c.base), int(batch), ws.data_ptr(), ws.numel(),
                                                     ids.ctypes.data if ids.size else None, int(ids.size), _stream()))


def vbpr_project(cfg: VbprCfg, st, F):
    """Refresh st['V'][:, k/2:] = F.E and st['bsum'] = rb + F.c."""
    f32 = torch.float32
    _need_cuda(F, st["V"])
    with torch.cuda.device(F.device):
        _check(lib().tkr_vbpr_project(cfg.ptr, _dev(F, f32, "F"), _dev(st["E"], f32, "E"), _dev(st["c"], f32, "c"), _dev(st["rb"], f32, "rb"),
                                      _dev(st["V"], f32, "V"), _dev(st["bsum"], f32, "bsum"), _stream()))


def vbpr_step(cfg: VbprCfg, st, F, u, i, j, batch, n_steps, ws, loss=None, sampler=None, first_draw=0):
    """st: dict of CUDA tensors U, V, rb, bsum, E, c and the slots msU, msV, msrb, msE, msc."""
    f32, i32 = torch.float32, torch.int32
    _need_cuda(F, st["U"], ws)
    with torch.cuda.device(F.device):
        _check(lib().tkr_vbpr_step(cfg.ptr, *(_dev(st[n], f32, n) for n in VBPR_STATE), _dev(F, f32, "F"),
                                   *(_dev(st.get(n), f32, n) for n in VBPR_SLOTS), _dev(u, i32, "u"), _dev(i, i32, "i"), _dev(j, i32, "j"),
                                   int(batch), int(n_steps), sampler.ptr if sampler is not None else None, int(first_draw),
                                   _dev(loss, f32, "loss"), ws.data_ptr(), ws.numel(), _stream()))


def vbpr_grad(cfg: VbprCfg, st, F, u, i, j, batch, ws, loss=None, sampler=None, first_draw=0, data_parallel=False):
    """first half of a data-parallel VBPR step: projection, gather/scatter gradients, this rank's dE / dc (tkr_vbpr_grad)"""
    f32, i32 = torch.float32, torch.int32
    _need_cuda(F, st["U"], ws)
    with torch.cuda.device(F.device):
        _check(lib().tkr_vbpr_grad(cfg.ptr, *(_dev(st[n], f32, n) for n in VBPR_STATE), _dev(F, f32, "F"),
                                   _dev(u, i32, "u"), _dev(i, i32, "i"), _dev(j, i32, "j"), int(batch),
                                   sampler.ptr if sampler is not None else None, int(first_draw), _dev(loss, f32, "loss"),
                                   ws.data_ptr(), ws.numel(), int(bool(data_parallel)), _stream()))


def vbpr_apply(cfg: VbprCfg, st, batch, ws, loss=None, data_parallel=False):
    """second half: sparse + dense optimiser updates from the (summed) gradients in the workspace (tkr_vbpr_apply)"""
    f32 = torch.float32
    _need_cuda(st["U"], ws)
    with torch.cuda.device(ws.device):
        _check(lib().tkr_vbpr_apply(cfg.ptr, *(_dev(st[n], f32, n) for n in ("U", "V", "rb", "E", "c")),
                                    *(_dev(st.get(n), f32, n) for n in VBPR_SLOTS), int(batch), _dev(loss, f32, "loss"),
                                    ws.data_ptr(), ws.numel(), int(bool(data_parallel)), _stream()))


def vbpr_grad_views(cfg: VbprCfg, batch, ws):
    """fp32 views of the two workspace regions a data-parallel caller sums over the ranks: [GV|Gb|tchV] and [GE|Gc]"""
    off = (C.c_int64 * 4)()
    _check(lib().tkr_vbpr_workspace_layout(cfg.ptr, int(batch), off))
    return ws[off[0]:off[1]].view(torch.float32), ws[off[2]:off[3]].view(torch.float32)


def bpr_sample(sampler: Sampler, first_draw, n, device="cuda"):
    _need_cuda()
    u = torch.empty(n, dtype=torch.int32, device=device)
    i = torch.empty_like(u); j = torch.empty_like(u)
    with torch.cuda.device(u.device):
        _check(lib().tkr_bpr_sample(sampler.ptr, int(first_draw), int(n), u.data_ptr(), i.data_ptr(), j.data_ptr(), _stream()))
    return u, i, j


def score_topk(U, V, k, bias=None, rated_indptr=None, rated_idx=None, col_offset=0, out=None, ws=None, engine="exact",
               n_fallback=None, items_prepared=False):
    """Device tensors in, device tensors out: (idx int32 [nu,k], score fp32 [nu,k]).
    engine='exact': fp32 CUDA-core kernel; engine='tc': tcgen05 BF16 filter + exact refine (same bits).
    n_fallback (tc only): optional int32 CUDA tensor [1] receiving the number of rows the exact kernel re-did.
    items_prepared (tc only): reuse the BF16 item table the previous call left in `ws` (same V, bias, shapes)."""
    f32 = torch.float32
    _need_cuda(U, V)
    nu, d = U.shape
    ni = V.shape[0]
    if V.shape[1] != d:
        raise ValueError("U and V disagree on d")
    if engine not in ("exact", "tc"):
        raise ValueError("engine must be 'exact' or 'tc'")
    if out is None:
        out = (torch.empty((nu, k), dtype=torch.int32, device=U.device), torch.empty((nu, k), dtype=f32, device=U.device))
    if engine == "tc":
        need = lib().tkr_score_topk_tc_workspace_bytes(nu, ni, d, k, int(bias is not None))
    else:
        need = lib().tkr_score_topk_workspace_bytes(nu, ni, d, k)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=U.device)
    if rated_indptr is not None and (rated_idx is None or rated_idx.numel() == 0):
        rated_idx = torch.zeros(1, dtype=torch.int32, device=U.device)
    args = (_dev(U, f32, "U"), nu, _dev(V, f32, "V"), ni, d, _dev(bias, f32, "bias"),
            _dev(rated_indptr, torch.int64, "rated_indptr"), _dev(rated_idx, torch.int32, "rated_idx"),
            int(k), int(col_offset), out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(), ws.numel())
    with torch.cuda.device(U.device):
        if engine == "tc":
            _check(lib().tkr_score_topk_tc(*args, _dev(n_fallback, torch.int32, "n_fallback"), int(bool(items_prepared)), _stream()))
        else:
            _check(lib().tkr_score_topk(*args, _stream()))
    return out


def score_topk_segment(U, V_shard, k, col_offset, state, first, last, V_full=None, bias_shard=None, bias_full=None, rated_indptr=None,
                       rated_idx=None, out=None, ws=None, n_fallback=None, items_prepared=False):
    """One segment of a tensor-core sweep over an item table cut into shards (tkr_score_topk_tc_segment): ``state`` (uint8 CUDA
    tensor of tkr_score_topk_tc_state_bytes(nu)) carries thresholds and candidates from segment to segment; the last segment
    returns (idx, score) for the whole table, bit-identical to ``score_topk`` on it."""
    f32 = torch.float32
    _need_cuda(U, V_shard, state)
    nu, d = U.shape
    ni = V_shard.shape[0]
    ni_full = V_full.shape[0] if V_full is not None else 0
    if last and out is None:
        out = (torch.empty((nu, k), dtype=torch.int32, device=U.device), torch.empty((nu, k), dtype=f32, device=U.device))
    need = lib().tkr_score_topk_tc_segment_workspace_bytes(nu, ni, max(ni_full, ni), d, k, int(bias_shard is not None))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=U.device)
    if rated_indptr is not None and (rated_idx is None or rated_idx.numel() == 0):
        rated_idx = torch.zeros(1, dtype=torch.int32, device=U.device)
    with torch.cuda.device(U.device):
        _check(lib().tkr_score_topk_tc_segment(_dev(U, f32, "U"), nu, _dev(V_shard, f32, "V_shard"), ni, d, _dev(bias_shard, f32, "bias_shard"),
                                               _dev(rated_indptr, torch.int64, "rated_indptr"), _dev(rated_idx, torch.int32, "rated_idx"), int(k),
                                               int(col_offset), state.data_ptr(), int(bool(first)), int(bool(last)), _dev(V_full, f32, "V_full"), ni_full,
                                               _dev(bias_full, f32, "bias_full"), out[0].data_ptr() if out else None, out[1].data_ptr() if out else None,
                                               ws.data_ptr(), ws.numel(), _dev(n_fallback, torch.int32, "n_fallback"), int(bool(items_prepared)), _stream()))
    return out


def score_topk_host(U, V, k, bias=None, rated_indptr=None, rated_idx=None, dev=None, device="cuda"):
    """numpy (host) arrays in and out; H2D/D2H copies happen inside the call
    (the np.dot + np.argsort seam of evaluate.py:78-81)."""
    _need_cuda()
    U = np.ascontiguousarray(U, np.float32); V = np.ascontiguousarray(V, np.float32)
    nu, d = U.shape; ni = V.shape[0]
    if bias is not None:
        bias = np.ascontiguousarray(bias, np.float32).ravel()
    n_rated = 0
    if rated_indptr is not None:
        rated_indptr = np.ascontiguousarray(rated_indptr, np.int64)
        rated_idx = np.ascontiguousarray(rated_idx if rated_idx is not None else np.zeros(0), np.int32)
        n_rated = int(rated_indptr[-1])
    need = lib().tkr_score_topk_host_device_bytes(nu, ni, d, k, n_rated)
    if dev is None or dev.numel() < need:
        dev = torch.empty(need, dtype=torch.uint8, device=device)
    out_idx = np.empty((nu, k), np.int32); out_score = np.empty((nu, k), np.float32)
    p = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    with torch.cuda.device(dev.device):
        _check(lib().tkr_score_topk_host(p(U), nu, p(V), ni, d, p(bias), p(rated_indptr),
                                         p(rated_idx) if rated_indptr is not None else None,
                                         int(k), p(out_idx), p(out_score), dev.data_ptr(), dev.numel(), _stream()))
    return out_idx, out_score


class BatchScorer:
    """The evaluator's loop (evaluate.py:75-81 over all users) with the PCIe traffic hidden: a HOST float32 user matrix
    is scored ``user_batch`` rows at a time against a device item table; the upload of batch t+1 and the download of
    the lists of batch t-1 run on side streams while batch t computes.  Holds the streams, the double buffers and the
    engine workspace, so it can be reused across passes (same shapes)."""

    def __init__(self, n_items, d, k, user_batch=18944, has_bias=False, rated=False, engine="tc", device="cuda"):
        _need_cuda()
        self.dev = dev = torch.device(device)
        self.ni, self.d, self.k, self.nb, self.engine, self.has_bias = int(n_items), int(d), int(k), int(user_batch), engine, bool(has_bias)
        f32, i32 = torch.float32, torch.int32
        need = (lib().tkr_score_topk_tc_workspace_bytes(self.nb, self.ni, self.d, self.k, int(self.has_bias)) if engine == "tc"
                else lib().tkr_score_topk_workspace_bytes(self.nb, self.ni, self.d, self.k))
        self.ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.Ud = [torch.empty((self.nb, d), dtype=f32, device=dev) for _ in range(2)]
        self.gi = [torch.empty((self.nb, k), dtype=i32, device=dev) for _ in range(2)]
        self.gs = [torch.empty((self.nb, k), dtype=f32, device=dev) for _ in range(2)]
        self.rp = [torch.empty(self.nb + 1, dtype=torch.int64, device=dev) for _ in range(2)] if rated else None
        self.ev = [[torch.cuda.Event() for _ in range(2)] for _ in range(4)]      # up, free, out, down
        self._prepared_for = None

    def run(self, umat, V, bias=None, rated_indptr=None, rated_idx=None, out=None, to_host=True):
        """umat: host float32 [n, d] (numpy or torch; pinned for real overlap).  rated_indptr: host int64 [n+1] over the
        device CSR ``rated_idx``.  Returns (idx, score) for all users: pinned host tensors when ``to_host`` (or the
        preallocated ``out``), else device tensors."""
        _need_cuda(V)
        dev, nb, k = self.dev, self.nb, self.k
        U_h = umat if isinstance(umat, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(umat, np.float32))
        n = U_h.shape[0]
        if U_h.shape[1] != self.d or V.shape != (self.ni, self.d) or (bias is not None) != self.has_bias or (rated_indptr is not None) != (self.rp is not None):
            raise ValueError("BatchScorer was built for other shapes / options")
        if out is None:
            out = ((torch.empty((n, k), dtype=torch.int32).pin_memory(), torch.empty((n, k), dtype=torch.float32).pin_memory()) if to_host
                   else (torch.empty((n, k), dtype=torch.int32, device=dev), torch.empty((n, k), dtype=torch.float32, device=dev)))
        rp_h = torch.from_numpy(np.ascontiguousarray(rated_indptr, np.int64)) if rated_indptr is not None else None
        main = torch.cuda.current_stream(dev)
        up, down = self.up, self.down
        ev_up, ev_free, ev_out, ev_down = self.ev
        starts = list(range(0, n, nb))

        def upload(t):
            r0 = starts[t]; r1 = min(n, r0 + nb); s = t & 1
            with torch.cuda.stream(up):
                if t >= 2:
                    up.wait_event(ev_free[s])              # batch t-2 no longer reads this slot
                self.Ud[s][:r1 - r0].copy_(U_h[r0:r1], non_blocking=True)
                if rp_h is not None:
                    self.rp[s][:r1 - r0 + 1].copy_(rp_h[r0:r1 + 1], non_blocking=True)
                ev_up[s].record(up)

        up.wait_stream(main); down.wait_stream(main)
        if starts:
            upload(0)
        for t, r0 in enumerate(starts):
            r1 = min(n, r0 + nb); m = r1 - r0; s = t & 1
            if t + 1 < len(starts):
                upload(t + 1)
            main.wait_event(ev_up[s])
            if t >= 2:
                main.wait_event(ev_down[s])                # the lists of batch t-2 have left this slot
            score_topk(self.Ud[s][:m], V, k, bias, self.rp[s][:m + 1] if self.rp is not None else None, rated_idx, engine=self.engine,
                       ws=self.ws, out=(self.gi[s][:m], self.gs[s][:m]), items_prepared=(self.engine == "tc" and t > 0 and m == nb))
            ev_free[s].record(main); ev_out[s].record(main)
            with torch.cuda.stream(down):
                down.wait_event(ev_out[s])
                out[0][r0:r1].copy_(self.gi[s][:m], non_blocking=True)
                out[1][r0:r1].copy_(self.gs[s][:m], non_blocking=True)
                ev_down[s].record(down)
        main.wait_stream(down); main.wait_stream(up)
        main.synchronize()
        return out


def score_topk_batches(umat, V, k, bias=None, rated_indptr=None, rated_idx=None, user_batch=18944, engine="tc", to_host=True):
    """One pass of ``BatchScorer`` (built for this call)."""
    n = umat.shape[0]
    sc = BatchScorer(V.shape[0], V.shape[1], k, min(int(user_batch), n), bias is not None, rated_indptr is not None, engine, V.device)
    return sc.run(umat, V, bias, rated_indptr, rated_idx, to_host=to_host)


def topk_merge(idx, score, out=None):
    """idx/score: device tensors [n_lists, nu, k] -> merged [nu, k]."""
    _need_cuda(idx, score)
    n_lists, nu, k = idx.shape
    if out is None:
        out = (torch.empty((nu, k), dtype=torch.int32, device=idx.device), torch.empty((nu, k), dtype=torch.float32, device=idx.device))
    with torch.cuda.device(idx.device):
        _check(lib().tkr_topk_merge(_dev(idx, torch.int32, "idx"), _dev(score, torch.float32, "score"), n_lists, nu, k,
                                    out[0].data_ptr(), out[1].data_ptr(), _stream()))
    return out


def eval_hits(lists, line_rows, likes_indptr, likes_idx, step, pos_hits=None):
    """Device hit counting (evaluate.py:84-112).  lists: int32 CUDA [n_rows, total]; line_rows int32 [n_lines];
    likes_indptr int64 [n_lines+1]; likes_idx int32 (ascending, distinct per line).  Returns the reference's
    cumulative hits[total // step] as float64 numpy (and the uint64 per-position histogram tensor)."""
    _need_cuda(lists, line_rows, likes_indptr, likes_idx)
    total = lists.shape[1]
    n_lines = line_rows.numel()
    if pos_hits is None:
        pos_hits = torch.zeros(total, dtype=torch.int64, device=lists.device)
    if likes_idx.numel() == 0:
        likes_idx = torch.zeros(1, dtype=torch.int32, device=lists.device)
    with torch.cuda.device(lists.device):
        _check(lib().tkr_eval_hits(_dev(lists, torch.int32, "lists"), int(total), _dev(line_rows, torch.int32, "line_rows"),
                                   _dev(likes_indptr, torch.int64, "likes_indptr"), _dev(likes_idx, torch.int32, "likes_idx"),
                                   int(n_lines), pos_hits.data_ptr(), _stream()))
    ph = pos_hits.cpu().numpy().astype(np.float64)
    interval = total // step
    return np.array([ph[:(q + 1) * step].sum() for q in range(interval)]), pos_hits


# ------------------------------------------------------------------ text codecs (host memory, no GPU needed)
def dat_read(path):
    """``.dat`` text matrix -> float32 ndarray (same roundings as ``np.float32(token)``; utils.py:28-44)."""
    rows, cols = C.c_int64(), C.c_int64()
    _check(lib().tkr_dat_shape(os.fsencode(path), C.byref(rows), C.byref(cols)))
    out = np.empty((rows.value, cols.value), np.float32)
    _check(lib().tkr_dat_read(os.fsencode(path), out.ctypes.data, rows.value, cols.value))
    return out


def dat_write(path, mat):
    """matrix -> ``.dat`` text, byte-identical to the reference writer (utils.py:47-55): float64 input is formatted from
    the doubles (what ``'%f ' % x`` does for CER's float64 ``E``), anything else as float32."""
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError("embed must be a matrix, got shape %s" % (mat.shape,))
    if mat.dtype == np.float64:
        mat = np.ascontiguousarray(mat)
        _check(lib().tkr_dat_write_f64(os.fsencode(path), mat.ctypes.data, mat.shape[0], mat.shape[1]))
        return
    mat = np.ascontiguousarray(mat, np.float32)
    _check(lib().tkr_dat_write(os.fsencode(path), mat.ctypes.data, mat.shape[0], mat.shape[1]))


def ratings_parse(ratings_path, uid_path, iid_path):
    """Rating file -> (line_user int32[L], line_indptr int64[L+1], pair_item int32[P], pair_like int8[P]); ids are rows
    of the two id files, -1 when unknown (utils.py:58-89, evaluate.py:30-45)."""
    nl, npair = C.c_int64(), C.c_int64()
    paths = [os.fsencode(p) for p in (ratings_path, uid_path, iid_path)]
    _check(lib().tkr_ratings_parse(*paths, C.byref(nl), C.byref(npair), None, None, None, None))
    line_user = np.empty(nl.value, np.int32); indptr = np.empty(nl.value + 1, np.int64)
    item = np.empty(npair.value, np.int32); like = np.empty(npair.value, np.int8)
    _check(lib().tkr_ratings_parse(*paths, C.byref(nl), C.byref(npair), line_user.ctypes.data, indptr.ctypes.data,
                                   item.ctypes.data, like.ctypes.data))
    return line_user, indptr, item, like

"""Peer-mapped exchange buffers (``tkr_peer_*``, include/topkrec.h): one process per GPU on one box; every rank
allocates a buffer of the same size, the CUDA-IPC handles travel over ``torch.distributed`` and every rank maps the
others' buffers, so the fused multi-GPU kernels can load from / store to peer HBM over NVLink.  torch is only the
rendezvous (one ``all_gather_object`` of 64-byte handles) and the carrier of the local view."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import lib, _check, _need_cuda

MAX_PEERS = 8


class tkr_peers(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("base", C.c_void_p * MAX_PEERS)]


class _RawCuda:
    """exposes a raw device allocation through the CUDA array interface so torch can alias it"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _bind():
    L = lib()
    if getattr(L, "_peer_bound", False):
        return L
    vp, sz = C.c_void_p, C.c_size_t
    L.tkr_peer_alloc.argtypes = [sz, C.POINTER(vp)]
    L.tkr_peer_free.argtypes = [vp]
    L.tkr_peer_export.argtypes = [vp, vp]
    L.tkr_peer_import.argtypes = [vp, C.POINTER(vp)]
    L.tkr_peer_release.argtypes = [vp]
    for n in ("tkr_peer_alloc", "tkr_peer_free", "tkr_peer_export", "tkr_peer_import", "tkr_peer_release"):
        getattr(L, n).restype = C.c_int
    L._peer_bound = True
    return L


class PeerBuffer:
    """``nbytes`` of zero-filled device memory on this rank's current device, mapped by every rank of ``group``.
    ``.local`` is a uint8 torch view of the own allocation, ``.peers`` the ``tkr_peers`` table for the kernels.
    Collective: every rank of the group must construct it (same ``nbytes``)."""

    def __init__(self, nbytes, group=None, device=None):
        _need_cuda()
        L = _bind()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > MAX_PEERS:
            raise ValueError("at most %d ranks (one box) can share peer buffers" % MAX_PEERS)
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.nbytes = int(nbytes)
        self._imported = []
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _check(L.tkr_peer_alloc(self.nbytes, C.byref(p)))
            self.ptr = p.value
            handle = (C.c_ubyte * 64)()
            _check(L.tkr_peer_export(self.ptr, handle))
            ptrs = [None] * self.world
            ptrs[self.rank] = self.ptr
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    q = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    _check(L.tkr_peer_import(buf, C.byref(q)))
                    ptrs[r] = q.value
                    self._imported.append(q.value)
        self.ptrs = ptrs
        self.peers = tkr_peers(self.rank, self.world, (C.c_void_p * MAX_PEERS)(*(ptrs + [None] * (MAX_PEERS - self.world))))
        self.local = torch.as_tensor(_RawCuda(self.ptr, self.nbytes), device=self.device)

    @property
    def peers_ptr(self):
        return C.byref(self.peers)

    def view(self, offset, shape, dtype=torch.float32):
        """typed view of the LOCAL buffer at a byte offset"""
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty(0, dtype=dtype).element_size()
        return self.local[offset:offset + nb].view(dtype).view(*shape)

    def close(self):
        """collective in spirit: call on every rank once no kernel uses the buffers any more"""
        L = _bind()
        if self.ptr is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)         # nobody may still be reading what is about to be unmapped / freed
            for q in self._imported:
                L.tkr_peer_release(q)
            self._imported = []
            self.local = None
            if self.world > 1:
                dist.barrier(group=self.group)
            L.tkr_peer_free(self.ptr)
            self.ptr = None

"""Packed device copy of one 4-layer SELU MLP (idff_weights_create in include/idff_b200.h)."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib


class PackedWeights:
    """Owns an idff_weights_t*.  Built from the Linear layers of a reference-shaped network
    (torchcfm/models/models.py:9-17) and, for MLP_Embedding, its embedding table (:30)."""

    def __init__(self, Ws, bs, emb=None, device=None):
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise _lib.IdffError("PackedWeights needs a CUDA device: idff_b200 has no CPU path")
        self.device = device
        Ws = [np.ascontiguousarray(np.asarray(w, dtype=np.float32)) for w in Ws]
        bs = [np.ascontiguousarray(np.asarray(b, dtype=np.float32)) for b in bs]
        if len(Ws) != 4 or len(bs) != 4:
            raise _lib.IdffError("expected 4 Linear layers (Linear-SELU x3 - Linear)")
        dims = [Ws[0].shape[1]] + [w.shape[0] for w in Ws]
        emb_rows = emb_dim = 0
        emb_p = None
        if emb is not None:
            emb = np.ascontiguousarray(np.asarray(emb, dtype=np.float32))
            emb_rows, emb_dim = emb.shape
            emb_p = emb.ctypes.data_as(C.c_void_p)
        self.dims = dims
        self.emb_rows, self.emb_dim = emb_rows, emb_dim
        self.din = dims[0] - emb_dim
        self.dout = dims[4]
        self.width = dims[1]
        wp = (C.c_void_p * 4)(*[w.ctypes.data_as(C.c_void_p) for w in Ws])
        bp = (C.c_void_p * 4)(*[b.ctypes.data_as(C.c_void_p) for b in bs])
        dm = (C.c_int32 * 5)(*dims)
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().idff_weights_create(wp, bp, dm, 4, emb_p, emb_rows, emb_dim, idx, C.byref(h)),
                   "idff_weights_create")
        self.handle = h
        self._fin = weakref.finalize(self, _destroy, h.value)

    @classmethod
    def from_module(cls, net, device=None):
        """net: an nn.Module whose `.net` is Sequential(Linear, SELU, Linear, SELU, Linear, SELU, Linear),
        optionally with `.time_index_embedding` (MLP_Embedding)."""
        lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
        act = [m for m in net.net if not isinstance(m, torch.nn.Linear)]
        if len(lin) != 4 or not all(isinstance(a, torch.nn.SELU) for a in act):
            raise _lib.IdffError("the fused field needs Linear-SELU-Linear-SELU-Linear-SELU-Linear")
        emb = getattr(net, "time_index_embedding", None)
        if device is None:
            device = lin[0].weight.device if lin[0].weight.is_cuda else "cuda"
        return cls([m.weight.detach().cpu().numpy() for m in lin], [m.bias.detach().cpu().numpy() for m in lin],
                   None if emb is None else emb.weight.detach().cpu().numpy(), device=device)

    @classmethod
    def from_npz(cls, path_or_dict, device=None):
        d = np.load(path_or_dict) if isinstance(path_or_dict, str) else path_or_dict
        emb = d["emb"] if "emb" in getattr(d, "files", d) else None
        return cls([d[f"W{i}"] for i in range(4)], [d[f"b{i}"] for i in range(4)], emb, device=device)


def _destroy(h):
    try:
        _lib.lib().idff_weights_destroy(C.c_void_p(h))
    except Exception:
        pass


_CACHE = weakref.WeakKeyDictionary()


def packed_for(net, device) -> PackedWeights:
    """Cached PackedWeights of `net` on `device`; rebuilt when a parameter was modified in place or replaced."""
    key = tuple((p.data_ptr(), p._version) for p in net.parameters())
    ent = _CACHE.get(net)
    dev = torch.device(device)
    if ent is not None and ent[0] == key and ent[1].device == dev:
        return ent[1]
    pw = PackedWeights.from_module(net, dev)
    _CACHE[net] = (key, pw)
    return pw

"""Glue between the torchcfm-shaped Python surface and the C-ABI."""
import ctypes as C

import torch

from .. import _lib
from ..weights import packed_for


def use_kernel(module, x):
    """The CUDA kernel serves inference calls: CUDA fp32 input, autograd off or nothing requiring grad."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        return False
    return True


def mlp_forward(module, x, t_idx):
    pw = packed_for(module, x.device)
    if x.shape[1] != pw.din:
        raise _lib.IdffError(f"MLP input has {x.shape[1]} columns, network expects {pw.din}")
    x = x.contiguous()
    out = torch.empty((x.shape[0], pw.dout), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().idff_mlp_forward(pw.handle, _lib.ptr(x), _lib.ptr(out), x.shape[0], int(t_idx),
                                           _lib.stream_of(x)), "idff_mlp_forward")
    return out

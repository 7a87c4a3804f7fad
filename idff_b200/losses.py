"""Fused training losses (forward value + gradient wrt the network output in one kernel pass), exposed as
autograd functions so `loss.backward()` continues into the torch MLP.

    flow_loss(vt, x1, t)                 mean(((vt-x1)/(1e-2+1-t))^2)        Autograd_example_order2.py:105
    score_loss(st, xt, x1, x0, t, s0sq)  mean((st - log p_t(xt|x1))^2)       example_order1_with_prediction_head.py:26-51,124-132
    md_periodic_loss(vt, x1)             sum_k mean((vt-(x1+k))^2), k=0,+-360  MD_simulation.py:137-141
"""
import torch

from . import _lib

_scratch = {}


def _scratch_for(device):
    key = (device.type, device.index)
    buf = _scratch.get(key)
    if buf is None:
        buf = torch.zeros(int(_lib.lib().idff_loss_scratch_bytes()), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _run(kind, a, x1, x0, xt, t, sigma0_sq, want_grad):
    for n, v in (("a", a), ("x1", x1)):
        _lib.require_cuda(v, n)
    ac, x1c = a.detach().contiguous(), x1.detach().contiguous()
    B = ac.shape[0]
    D = ac.numel() // max(B, 1)
    x0c = None if x0 is None else x0.detach().contiguous()
    xtc = None if xt is None else xt.detach().contiguous()
    tc = None if t is None else t.detach().contiguous()
    loss = torch.empty((), device=ac.device, dtype=torch.float32)
    grad = torch.empty_like(ac) if want_grad else None
    _lib.check(_lib.lib().idff_loss(kind, _lib.ptr(ac), _lib.ptr(x1c), _lib.ptr(x0c), _lib.ptr(xtc), _lib.ptr(tc),
                                    float(sigma0_sq), 1.0, _lib.ptr(loss), _lib.ptr(grad), B, D,
                                    _lib.ptr(_scratch_for(ac.device)), _lib.stream_of(ac)), "idff_loss")
    return loss, grad


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, sigma0_sq, a, x1, x0, xt, t):
        loss, grad = _run(kind, a, x1, x0, xt, t, sigma0_sq, ctx.needs_input_grad[2])
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, None, (grad * g if grad is not None else None), None, None, None, None


def flow_loss(vt, x1, t):
    return _Loss.apply(0, 0.0, vt, x1, None, None, t)


def score_loss(st, xt, x1, x0, t, sigma0_sq):
    return _Loss.apply(1, float(sigma0_sq), st, x1, x0, xt, t)


def md_periodic_loss(vt, x1):
    return _Loss.apply(2, 0.0, vt, x1, None, None, None)

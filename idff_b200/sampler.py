"""Integrator drivers with the third-party call shapes the reference uses:

    odeint(model, x0, t_span, rtol=.., atol=.., method='rk4')     torchdiffeq, Autograd_example_order2.py:167,213,291
    sdeint(sde, y0, ts=.., dt=..)                                 torchsde,    MD_simulation.py:209,228

When `model` / `sde` is one of the idff_b200.fields wrappers the WHOLE solve is one launch of the fused sampler
kernel (idff_sample_rk4 / idff_sample_sde): particle state stays on chip across all stages.  Any other callable is
stepped on the host with the same fixed-grid 3/8-rule scheme (4 field calls per interval)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .fields import SDE_cal, _FieldBase

_ONE_THIRD, _TWO_THIRDS = 1.0 / 3.0, 2.0 / 3.0


def sample_rk4(model: _FieldBase, x0, t_span, return_traj=False, out=None):
    """Fused fixed-grid RK4 (3/8 rule).  Returns traj[-1] (N, D), or the whole traj (len(t), N, D)."""
    _lib.require_cuda(x0, "x0")
    if x0.dim() != 2:
        raise _lib.IdffError(f"x0 must be (N, D), got {tuple(x0.shape)}")
    x0c = x0.detach().contiguous()
    tg = np.ascontiguousarray(t_span.detach().cpu().numpy() if torch.is_tensor(t_span) else np.asarray(t_span),
                              dtype=np.float32)
    N, D = x0c.shape
    pw = model._packed(x0c.device)
    p = model.field_params(D)
    final = out if out is not None else torch.empty_like(x0c)
    traj = torch.empty((len(tg), N, D), device=x0c.device, dtype=torch.float32) if return_traj else None
    _lib.check(_lib.lib().idff_sample_rk4(pw.handle, C.byref(p), _lib.ptr(x0c), tg.ctypes.data_as(C.c_void_p),
                                          len(tg), _lib.ptr(final), _lib.ptr(traj), N, _lib.stream_of(x0c)),
               "idff_sample_rk4")
    return traj if return_traj else final


def set_small_n_route(n_max: int = 4736):
    """Opt in (calling thread): AUTO solves of the order-1 / order-2 field with at most n_max particles run on 32-particle tiles
    (the three-jet kernel with an idle jet) -- 1.25 x faster up to 4736 particles on a B200; results then depend in their last bits
    on the batch size (see include/idff_b200.h).  0 turns it off (the default)."""
    _lib.check(_lib.lib().idff_tune_small_n_route(int(n_max)), "idff_tune_small_n_route")


def sample_rk4_sweep(models, x0, t_span, out=None):
    """The same x0 and grid solved under every field of `models` (wrappers around ONE packed network that differ in
    gammas / order / sigma): -> (len(models), N, D) = traj[-1] per field.  One launch for the whole sweep where the
    3xFP16 tensor kernel serves the fields (idff_sample_rk4_sweep), identical to per-field sample_rk4 calls."""
    if len(models) == 0:
        raise _lib.IdffError("sample_rk4_sweep: no fields given")
    _lib.require_cuda(x0, "x0")
    if x0.dim() != 2:
        raise _lib.IdffError(f"x0 must be (N, D), got {tuple(x0.shape)}")
    x0c = x0.detach().contiguous()
    tg = np.ascontiguousarray(t_span.detach().cpu().numpy() if torch.is_tensor(t_span) else np.asarray(t_span),
                              dtype=np.float32)
    N, D = x0c.shape
    pw = models[0]._packed(x0c.device)
    for m in models[1:]:
        if m._packed(x0c.device) is not pw:
            raise _lib.IdffError("sample_rk4_sweep: every field must wrap the same packed network")
    params = (_lib.FieldParams * len(models))(*[m.field_params(D) for m in models])
    res = out if out is not None else torch.empty((len(models), N, D), device=x0c.device, dtype=torch.float32)
    if tuple(res.shape) != (len(models), N, D) or not res.is_contiguous():
        raise _lib.IdffError(f"out must be a contiguous {(len(models), N, D)} tensor")
    _lib.check(_lib.lib().idff_sample_rk4_sweep(pw.handle, C.cast(params, C.c_void_p), len(models), _lib.ptr(x0c),
                                                tg.ctypes.data_as(C.c_void_p), len(tg), _lib.ptr(res), N,
                                                _lib.stream_of(x0c)), "idff_sample_rk4_sweep")
    return res


def odeint(func, y0, t, rtol=None, atol=None, method="rk4", options=None):
    """torchdiffeq.odeint for the fixed-grid 'rk4' method (rtol/atol are ignored by that solver). -> (len(t), *y0.shape)"""
    del rtol, atol, options
    if method != "rk4":
        raise _lib.IdffError(f"only the fixed-grid method='rk4' (3/8 rule) is provided, got {method!r}")
    if isinstance(func, _FieldBase) and not isinstance(func, SDE_cal):
        return sample_rk4(func, y0, t, return_traj=True)
    sol = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        k1 = func(t0, y)
        k2 = func(t0 + dt * _ONE_THIRD, y + dt * k1 * _ONE_THIRD)
        k3 = func(t0 + dt * _TWO_THIRDS, y + dt * (k2 - k1 * _ONE_THIRD))
        k4 = func(t1, y + dt * (k1 - k2 + k3))
        y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        sol.append(y)
    return torch.stack(sol)


def sde_num_steps(ts, dt) -> int:
    tg = np.ascontiguousarray(ts.detach().cpu().numpy() if torch.is_tensor(ts) else np.asarray(ts), dtype=np.float32)
    n = _lib.lib().idff_sde_num_steps(tg.ctypes.data_as(C.c_void_p), len(tg), float(dt))
    if n < 0:
        _lib.check(n, "idff_sde_num_steps")
    return n


def brownian_increments(n_steps, ts, dt, shape, device, generator=None):
    """(I_k, I_k0) for every solver step: W increments ~ N(0, h) and the space-time Levy term
    U = h/2 * W + N(0, h^3/12) (what torchsde's BrownianInterval returns with return_U=True)."""
    tg = np.asarray(ts.detach().cpu().numpy() if torch.is_tensor(ts) else ts, dtype=np.float32)
    grid = [np.float32(tg[0])]
    while grid[-1] < tg[-1]:
        grid.append(min(np.float32(grid[-1] + np.float32(dt)), np.float32(tg[-1])))
    h = np.diff(np.array(grid, dtype=np.float32))
    assert len(h) == n_steps
    I_k = torch.randn((n_steps,) + tuple(shape), device=device, generator=generator)
    I_k0 = torch.randn((n_steps,) + tuple(shape), device=device, generator=generator)
    # per-step scalings with host scalars (fp32 like the tensors): no host-to-device copy, so the draw can be captured
    # in a CUDA graph
    for s_ in range(n_steps):
        hs = np.float32(h[s_])
        I_k[s_].mul_(float(np.sqrt(hs)))
        I_k0[s_].mul_(float(hs ** np.float32(1.5)) / math.sqrt(12.0)).add_(I_k[s_], alpha=float(np.float32(0.5) * hs))
    return I_k, I_k0


def sdeint(sde, y0, ts, dt=1e-3, method="srk", bm=None, generator=None):
    """torchsde.sdeint(sde, y0, ts=ts, dt=dt) for SDE_cal: fixed-step srk (default for ito/diagonal) or euler.
    `bm` = (I_k, I_k0) injected Brownian increments of shape (n_steps, N, D); drawn on the device when omitted
    (the reference's own BrownianInterval is entropy-seeded, so its noise is not reproducible either).
    -> (len(ts), N, D), linearly interpolated at ts as torchsde does."""
    if not isinstance(sde, SDE_cal):
        raise _lib.IdffError("sdeint is provided for idff_b200.fields.SDE_cal objects")
    meth = {"srk": 0, "euler": 1}.get(method)
    if meth is None:
        raise _lib.IdffError(f"method must be 'srk' or 'euler', got {method!r}")
    _lib.require_cuda(y0, "y0")
    y0c = y0.detach().contiguous().view(-1, sde.input_dim)
    N, D = y0c.shape
    tg = np.ascontiguousarray(ts.detach().cpu().numpy() if torch.is_tensor(ts) else np.asarray(ts), dtype=np.float32)
    n_steps = sde_num_steps(tg, dt)
    if bm is None:
        bm = brownian_increments(n_steps, tg, dt, (N, D), y0c.device, generator)
    I_k, I_k0 = bm
    I_k = I_k.contiguous()
    I_k0 = None if I_k0 is None else I_k0.contiguous()
    if tuple(I_k.shape) != (n_steps, N, D):
        raise _lib.IdffError(f"I_k must be {(n_steps, N, D)}, got {tuple(I_k.shape)}")
    pw = sde._packed(y0c.device)
    p = sde.field_params(D)
    out = torch.empty((len(tg), N, D), device=y0c.device, dtype=torch.float32)
    _lib.check(_lib.lib().idff_sample_sde(pw.handle, C.byref(p), _lib.ptr(y0c), tg.ctypes.data_as(C.c_void_p), len(tg),
                                          float(dt), meth, _lib.ptr(I_k), _lib.ptr(I_k0), _lib.ptr(out), N,
                                          _lib.stream_of(y0c)), "idff_sample_sde")
    return out

"""The IDFF field wrappers with the reference's constructor signatures and the forward(t, x) / f(t, y), g(t, y)
call conventions, so they plug into odeint / sdeint unchanged.  One call = one launch of the B200 field kernel
(idff_field_eval); the higher-order terms the reference obtains with nested autograd.grad are computed inside the
kernel as forward Taylor jets plus one reverse sweep, so no autograd graph exists.

    CustomNetModel        2D_examples/Autograd_example_order2.py:36-76     (order 2)
    CustomNetModelOrder3  2D_examples/Autograd_example_order3.py:34-77     (order 3)
    CustomNetModelScoreHead  2D_examples/example_order1_with_prediction_head.py:64-96
    CFMModel              2D_examples/Autograd_example_order2.py:310-332
    SDE_cal               timeseris_example/MD_simulation.py:481-529
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .weights import PackedWeights, packed_for


def _t_float(t):
    return float(t.item()) if torch.is_tensor(t) else float(t)


class _FieldBase(torch.nn.Module):
    variant = _lib.IDFF_FIELD_MOMENTUM
    order = 2
    impl = _lib.IDFF_IMPL_AUTO

    def _packed(self, device) -> PackedWeights:
        if isinstance(self.base_model, PackedWeights):
            return self.base_model
        return packed_for(self.base_model, device)

    def _gammas(self):
        return (0.0, 0.0, 0.0)

    def field_params(self, dim: int) -> _lib.FieldParams:
        p = _lib.FieldParams()
        p.variant, p.order, p.dim, p.t_idx, p.impl = self.variant, self.order, int(dim), self._t_idx(), int(self.impl)
        p.sigma = float(self.sigma)
        g = self._gammas()
        for i in range(3):
            p.gamma[i] = float(g[i])
        p.end_div = float(self._end_div())
        return p

    def _t_idx(self):
        return 0

    def _end_div(self):
        return 1e-2

    def _eval(self, t, x, x0=None):
        _lib.require_cuda(x, "x")
        if x.dim() != 2:
            raise _lib.IdffError(f"x must be (N, D), got {tuple(x.shape)}")
        xc = x.detach().contiguous()
        pw = self._packed(x.device)
        p = self.field_params(xc.shape[1])
        out = torch.empty_like(xc)
        _lib.check(_lib.lib().idff_field_eval(pw.handle, C.byref(p), _t_float(t), _lib.ptr(xc),
                                              _lib.ptr(x0), _lib.ptr(out), xc.shape[0], _lib.stream_of(xc)),
                   "idff_field_eval")
        return out


class CustomNetModel(_FieldBase):
    """Order-2 IDFF field: sqrt(1-(g1+g2) s_t^2) (x1hat-x)/(1-t) + .5(2 g1-1) g_1 t(1-t) + g2 g_2 t(1-t)."""
    order = 2

    def __init__(self, base_model, sigma=0.2, gamma1=0.5, gamma2=0.5):
        super().__init__()
        self.base_model = base_model
        self.sigma, self.gamma1, self.gamma2 = sigma, gamma1, gamma2

    def _gammas(self):
        return (self.gamma1, self.gamma2, 0.0)

    def forward(self, t, xt):
        return self._eval(t, xt)


class CustomNetModelOrder3(_FieldBase):
    """Order-3 field (adds the gamma3 term); constructed as CustomNetModel(base, sigma, gamma1, gamma2, gamma3) in
    Autograd_example_order3.py:35."""
    order = 3

    def __init__(self, base_model, sigma=0.2, gamma1=0.5, gamma2=0.5, gamma3=0.5):
        super().__init__()
        self.base_model = base_model
        self.sigma, self.gamma1, self.gamma2, self.gamma3 = sigma, gamma1, gamma2, gamma3

    def _gammas(self):
        return (self.gamma1, self.gamma2, self.gamma3)

    def forward(self, t, xt):
        return self._eval(t, xt)


class CustomNetModelScoreHead(_FieldBase):
    """Order-1 field with a score head: the MLP maps cat[x, x, t] -> (x1hat | st)."""
    variant = _lib.IDFF_FIELD_SCOREHEAD
    order = 1

    def __init__(self, base_model, sigma=0.2, gamma1=0.5, gamma2=0.5):
        super().__init__()
        self.base_model = base_model
        self.sigma, self.gamma1, self.gamma2 = sigma, gamma1, gamma2

    def _gammas(self):
        return (self.gamma1, self.gamma2, 0.0)

    def forward(self, t, xt):
        return self._eval(t, xt)


class CFMModel(_FieldBase):
    """Plain-CFM comparator: x1hat(t, x) - x0, x0 = the x seen at t == 0."""
    variant = _lib.IDFF_FIELD_CFM
    order = 1

    def __init__(self, base_model, sigma=0.2, gamma1=0.5, gamma2=0.5):
        super().__init__()
        self.base_model = base_model
        self.sigma, self.gamma1, self.gamma2 = sigma, gamma1, gamma2
        self.x0 = None

    def forward(self, t, xt):
        if _t_float(t) == 0.0:
            self.x0 = xt.detach().contiguous()
        if self.x0 is None:
            raise _lib.IdffError("CFMModel.forward called at t != 0 before any call at t == 0")
        return self._eval(t, xt, self.x0)


class SDE_cal(_FieldBase):
    """PolyALA SDE: drift f (order-2 field with the MD sign convention, t == 1 divides by dt) and diagonal
    diffusion g = 180 sigma^2 sqrt(t(1-t))."""
    noise_type = "diagonal"
    sde_type = "ito"
    variant = _lib.IDFF_FIELD_MD
    order = 2

    def __init__(self, model, input_dim, sigma=1, init_ind=1, max_length=1, dt=0.1, gamma1=1, gamma2=1, device="cuda"):
        super().__init__()
        self.model = model
        self.base_model = model
        self.sigma = sigma
        self.init_ind = torch.tensor([init_ind])
        self.max_length, self.dt = max_length, dt
        self.gamma1, self.gamma2 = gamma1, gamma2
        self.device, self.input_dim = device, input_dim

    def _gammas(self):
        return (self.gamma1, self.gamma2, 0.0)

    def _t_idx(self):
        return int(self.init_ind.item())

    def _end_div(self):
        return self.dt

    def f(self, t, y):
        return self._eval(t, y.view(-1, self.input_dim))

    def g(self, t, y):
        return torch.ones_like(y) * 180 * (self.sigma ** 2) * torch.sqrt(t * (1 - t))

    def forward(self, t, y):
        return self.f(t, y)

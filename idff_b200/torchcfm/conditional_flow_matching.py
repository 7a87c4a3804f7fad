"""Conditional-flow-matching path samplers behind the reference's API
(torchcfm/conditional_flow_matching.py:12-312), computed by the fused B200 path-sample kernel.

    xt = t*x1 + (1-t)*x0 + sigma_t*eps ,   ut = x1 - x0

`sample_location_and_conditional_flow` keeps the reference's draw order (t from torch.rand on the CPU generator,
then eps = randn_like(x0)) and produces xt and ut in ONE kernel launch, bit-exact with the reference's fp32
operation order.  Inputs must be CUDA float32 tensors; there is no CPU path."""
import ctypes as C
import math
from typing import Union

import torch

from .. import _lib

__all__ = ["pad_t_like_x", "ConditionalFlowMatcher", "ExactOptimalTransportConditionalFlowMatcher",
           "TargetConditionalFlowMatcher", "SchrodingerBridgeConditionalFlowMatcher",
           "VariancePreservingConditionalFlowMatcher"]


def pad_t_like_x(t, x):
    """(bs,) -> (bs, 1, ..., 1) so t broadcasts against x; numbers pass through (reference :12-33)."""
    if isinstance(t, (float, int)):
        return t
    return t.reshape(-1, *([1] * (x.dim() - 1)))


def _path_sample(x0, x1, t, eps, sigma, sigma_mode, want_ut=True, idx0=None, idx1=None):
    for n, v in (("x0", x0), ("x1", x1), ("t", t)):
        _lib.require_cuda(v, n)
    if x0.shape != x1.shape:
        raise _lib.IdffError(f"x0 {tuple(x0.shape)} and x1 {tuple(x1.shape)} differ")
    B = x0.shape[0] if idx0 is None else idx0.shape[0]
    D = x0[0].numel() if x0.shape[0] > 0 else int(math.prod(x0.shape[1:]))
    x0c, x1c, tc = x0.contiguous(), x1.contiguous(), t.contiguous()
    epsc = None
    if eps is not None:
        _lib.require_cuda(eps, "eps")
        epsc = eps.contiguous()
    out_shape = (B,) + tuple(x0.shape[1:])
    xt = torch.empty(out_shape, device=x0.device, dtype=torch.float32)
    ut = torch.empty(out_shape, device=x0.device, dtype=torch.float32) if want_ut else None
    _lib.check(_lib.lib().idff_path_sample(_lib.ptr(x0c), _lib.ptr(x1c), _lib.ptr(tc), _lib.ptr(epsc),
                                           _lib.ptr(idx0), _lib.ptr(idx1), float(sigma), int(sigma_mode),
                                           _lib.ptr(xt), _lib.ptr(ut), B, D, _lib.stream_of(x0)),
               "idff_path_sample")
    return xt, ut


def _path_sample_philox(x0, x1, t, sigma, sigma_mode, want_eps):
    """xt, ut[, eps] with eps = torch.randn_like(x0) drawn inside the kernel from the torch CUDA generator's Philox stream:
    the same numbers, element for element, and the generator ends where randn_like would have left it.  Returns None when
    the fused draw does not apply (stream capture, > 2^31 elements, non-contiguous / non-fp32 inputs)."""
    if not (x0.is_cuda and x1.is_cuda and t.is_cuda and x0.dtype == torch.float32 and x1.dtype == torch.float32
            and t.dtype == torch.float32 and x0.is_contiguous() and x1.is_contiguous() and t.is_contiguous()):
        return None
    n = x0.numel()
    if n == 0 or n >= 2 ** 31 or x0.shape != x1.shape or torch.cuda.is_current_stream_capturing():
        return None
    gen = torch.cuda.default_generators[x0.device.index if x0.device.index is not None else torch.cuda.current_device()]
    st = gen.get_state()                              # ... | seed (8 bytes LE) | philox offset (8 bytes LE)
    raw = bytes(st[-16:].tolist())
    seed, offset = int.from_bytes(raw[:8], "little"), int.from_bytes(raw[8:], "little")
    B, D = x0.shape[0], n // x0.shape[0]
    xt, ut = torch.empty_like(x0), torch.empty_like(x0)
    eps = torch.empty_like(x0) if want_eps else None
    inc = C.c_uint64(0)
    _lib.check(_lib.lib().idff_path_sample_philox(_lib.ptr(x0), _lib.ptr(x1), _lib.ptr(t), seed, offset, float(sigma),
                                                  int(sigma_mode), _lib.ptr(xt), _lib.ptr(ut), _lib.ptr(eps), B, D,
                                                  C.byref(inc), _lib.stream_of(x0)), "idff_path_sample_philox")
    st = st.clone()
    st[-8:] = torch.tensor(list((offset + inc.value).to_bytes(8, "little")), dtype=torch.uint8)
    gen.set_state(st)                                 # book the draw on the generator, as ATen's randn kernel does
    return xt, ut, eps


class ConditionalFlowMatcher:
    """Independent CFM: N(t*x1 + (1-t)*x0, sigma) (reference :36-211)."""
    _sigma_mode = 0      # sigma_t = sigma
    _lambda_dynamic = 0  # 2*sigma_t**2/(sigma**2+1e-8)   (reference :210-211)

    def __init__(self, sigma: Union[float, int] = 0.0):
        self.sigma = sigma

    # -- pieces (kept for API parity; the fused call below does not go through them) -------------------------
    def compute_mu_t(self, x0, x1, t):
        xt, _ = _path_sample(x0, x1, t, None, 0.0, 0, want_ut=False)
        return xt

    def compute_sigma_t(self, t):
        del t
        return self.sigma
    compute_sigma_t._idff_sigma_mode = 0     # the kernel's sigma_mode that computes exactly this function

    def _kernel_sigma_mode(self):
        """sigma_mode of the path-sample kernel that reproduces type(self).compute_sigma_t, or None when a subclass
        overrides it with something the kernel does not know (the reference routes everything through compute_sigma_t)."""
        return getattr(type(self).compute_sigma_t, "_idff_sigma_mode", None)

    def sample_xt(self, x0, x1, t, epsilon):
        mode = self._kernel_sigma_mode()
        if mode is None or type(self).compute_mu_t is not ConditionalFlowMatcher.compute_mu_t:
            return self.compute_mu_t(x0, x1, t) + pad_t_like_x(self.compute_sigma_t(t), x0) * epsilon   # reference :119-123
        xt, _ = _path_sample(x0, x1, t, epsilon, self.sigma, mode, want_ut=False)
        return xt

    def compute_conditional_flow(self, x0, x1, t, xt):
        del t, xt
        return x1 - x0

    def sample_noise_like(self, x):
        return torch.randn_like(x)

    def _fusable(self):
        c = type(self)
        base = ConditionalFlowMatcher
        return (c.compute_mu_t is base.compute_mu_t and c.sample_xt is base.sample_xt
                and c.compute_conditional_flow is base.compute_conditional_flow and self._kernel_sigma_mode() is not None)

    def sample_location_and_conditional_flow(self, x0, x1, t=None, return_noise=False):
        if t is None:
            t = torch.rand(x0.shape[0]).type_as(x0)          # CPU generator first (reference :184)
        assert len(t) == x0.shape[0], "t has to have batch size dimension"
        if (self._fusable() and "sample_noise_like" not in self.__dict__
                and type(self).sample_noise_like is ConditionalFlowMatcher.sample_noise_like):
            # noise drawn inside the path-sample kernel: the very numbers randn_like(x0) would produce (same Philox stream,
            # same element layout), without the round trip of eps through HBM
            got = _path_sample_philox(x0, x1, t, self.sigma, self._kernel_sigma_mode(), return_noise)
            if got is not None:
                xt, ut, eps = got
                return (t, xt, ut, eps) if return_noise else (t, xt, ut)
        eps = self.sample_noise_like(x0)                     # then the noise (reference :187)
        if self._fusable():
            xt, ut = _path_sample(x0, x1, t, eps, self.sigma, self._kernel_sigma_mode())
        else:
            xt = self.sample_xt(x0, x1, t, eps)
            ut = self.compute_conditional_flow(x0, x1, t, xt)
        if return_noise:
            return t, xt, ut, eps
        return t, xt, ut

    def compute_lambda(self, t):
        mode = self._kernel_sigma_mode()
        if mode is None:     # a compute_sigma_t the kernel does not know: the reference's expression (:210-211 / _dynamic :213-214)
            sigma_t = self.compute_sigma_t(t)
            return 2 * (sigma_t if self._lambda_dynamic else sigma_t ** 2) / (self.sigma ** 2 + 1e-8)
        _lib.require_cuda(t, "t")
        tc = t.contiguous()
        out = torch.empty_like(tc)
        _lib.check(_lib.lib().idff_compute_lambda(_lib.ptr(tc), float(self.sigma), mode,
                                                  self._lambda_dynamic, _lib.ptr(out), tc.numel(),
                                                  _lib.stream_of(tc)), "idff_compute_lambda")
        return out


class ExactOptimalTransportConditionalFlowMatcher(ConditionalFlowMatcher):
    """The matcher every IDFF script uses (reference :214-312): sigma_t = sigma*sqrt(t(1-t)); the OT re-pairing is
    disabled in this file (reference :265) and the batch is remembered in .x0/.x1 for the loss
    (Autograd_example_order2.py:105)."""
    _sigma_mode = 1

    def __init__(self, sigma: Union[float, int] = 0.0):
        super().__init__(sigma)
        self.ot_sampler = None

    def compute_sigma_t(self, t):
        return self.sigma * torch.sqrt(t * (1 - t))
    compute_sigma_t._idff_sigma_mode = 1

    def sample_location_and_conditional_flow(self, x0, x1, t=None, return_noise=False):
        self.x1 = x1
        self.x0 = x0
        return super().sample_location_and_conditional_flow(x0, x1, t, return_noise)

    def guided_sample_location_and_conditional_flow(self, x0, x1, y0=None, y1=None, t=None, return_noise=False):
        out = ConditionalFlowMatcher.sample_location_and_conditional_flow(self, x0, x1, t, return_noise)
        if return_noise:
            t, xt, ut, eps = out
            return t, xt, ut, y0, y1, eps
        t, xt, ut = out
        return t, xt, ut, y0, y1


class TargetConditionalFlowMatcher(ConditionalFlowMatcher):
    """Lipman-style target path (reference :315-388).  No IDFF configuration uses it: plain torch ops."""

    def compute_mu_t(self, x0, x1, t):
        del x0
        return pad_t_like_x(t, x1) * x1

    def compute_sigma_t(self, t):
        return 1 - (1 - self.sigma) * t

    def sample_xt(self, x0, x1, t, epsilon):
        return self.compute_mu_t(x0, x1, t) + pad_t_like_x(self.compute_sigma_t(t), x0) * epsilon

    def compute_conditional_flow(self, x0, x1, t, xt):
        del x0
        tp = pad_t_like_x(t, x1)
        return (x1 - (1 - self.sigma) * xt) / (1 - (1 - self.sigma) * tp)


class SchrodingerBridgeConditionalFlowMatcher(ConditionalFlowMatcher):
    """SB-CFM (reference :391-547): sigma_t = sigma sqrt(t(1-t)) (the path-sample kernel's mode 1) and the conditional flow
    (1-2t) / (2t(1-t) + 1e-8) (xt - mu_t) + x1 - x0.  In this reference file the OT re-pairing is commented out (:498, :533), so
    the matcher pairs the batches as given.  No IDFF configuration uses it: the flow is plain torch ops on the kernel's xt."""
    _sigma_mode = 1

    def __init__(self, sigma: Union[float, int] = 1.0, ot_method="exact"):
        if sigma <= 0:
            raise ValueError(f"Sigma must be strictly positive, got {sigma}.")
        if sigma < 1e-3:
            import warnings
            warnings.warn("Small sigma values may lead to numerical instability.")
        super().__init__(sigma)
        self.ot_method = ot_method
        self.ot_sampler = None      # never consulted by the reference's code path

    def compute_sigma_t(self, t):
        return self.sigma * torch.sqrt(t * (1 - t))
    compute_sigma_t._idff_sigma_mode = 1

    def compute_conditional_flow(self, x0, x1, t, xt):
        tp = pad_t_like_x(t, x0)
        mu_t = self.compute_mu_t(x0, x1, t)
        return (1 - 2 * tp) / (2 * tp * (1 - tp) + 1e-8) * (xt - mu_t) + x1 - x0

    def guided_sample_location_and_conditional_flow(self, x0, x1, y0=None, y1=None, t=None, return_noise=False):
        out = self.sample_location_and_conditional_flow(x0, x1, t, return_noise)
        if return_noise:
            t, xt, ut, eps = out
            return t, xt, ut, y0, y1, eps
        t, xt, ut = out
        return t, xt, ut, y0, y1


class VariancePreservingConditionalFlowMatcher(ConditionalFlowMatcher):
    """Trigonometric interpolant (reference :550-608).  No IDFF configuration uses it: plain torch ops."""

    def compute_mu_t(self, x0, x1, t):
        tp = pad_t_like_x(t, x0)
        return torch.cos(math.pi / 2 * tp) * x0 + torch.sin(math.pi / 2 * tp) * x1

    def sample_xt(self, x0, x1, t, epsilon):
        return self.compute_mu_t(x0, x1, t) + pad_t_like_x(self.compute_sigma_t(t), x0) * epsilon

    def compute_conditional_flow(self, x0, x1, t, xt):
        del xt
        tp = pad_t_like_x(t, x0)
        return math.pi / 2 * (torch.cos(math.pi / 2 * tp) * x1 - torch.sin(math.pi / 2 * tp) * x0)

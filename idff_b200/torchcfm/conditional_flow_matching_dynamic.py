"""The `_dynamic` flavour of the matchers (torchcfm/conditional_flow_matching_dynamic.py): identical to
conditional_flow_matching.py except that (a) compute_lambda returns 2*sigma_t/(sigma^2+1e-8) (:213-214) and
(b) the ExactOT matcher re-pairs the batch with the minibatch-OT plan before sampling (:268).  The re-pairing is
an index gather fused into the path-sample kernel; the plan comes from an OTPlanSampler -- on CUDA tensors the exact
assignment kernel idff_ot_assign, index pairs drawn on the device (no host round trip)."""
from typing import Union

import torch

from . import conditional_flow_matching as _cfm
from .conditional_flow_matching import pad_t_like_x  # noqa: F401

__all__ = ["pad_t_like_x", "ConditionalFlowMatcher", "ExactOptimalTransportConditionalFlowMatcher"]


class ConditionalFlowMatcher(_cfm.ConditionalFlowMatcher):
    _lambda_dynamic = 1


class ExactOptimalTransportConditionalFlowMatcher(ConditionalFlowMatcher):
    _sigma_mode = 1

    def __init__(self, sigma: Union[float, int] = 0.0, ot_sampler=None):
        super().__init__(sigma)
        if ot_sampler is None:
            from .optimal_transport import OTPlanSampler
            ot_sampler = OTPlanSampler(method="exact")
        self.ot_sampler = ot_sampler

    def compute_sigma_t(self, t):
        return self.sigma * torch.sqrt(t * (1 - t))
    compute_sigma_t._idff_sigma_mode = 1

    def _paired_path_sample(self, x0, x1, t, eps, i, j):
        mode = self._kernel_sigma_mode()
        if mode is None:     # a subclass changed compute_sigma_t: gather in torch, then the reference's expressions
            xi, xj = x0[i], x1[j]
            xt = self.sample_xt(xi, xj, t, eps)
            return xt, self.compute_conditional_flow(xi, xj, t, xt)
        return _cfm._path_sample(x0, x1, t, eps, self.sigma, mode, idx0=i, idx1=j)   # gather fused into the kernel

    def sample_location_and_conditional_flow(self, x0, x1, t=None, return_noise=False):
        i, j = self.ot_sampler.sample_map_indices(x0, x1)          # int64 device tensors, (bs,)
        if t is None:
            t = torch.rand(x0.shape[0]).type_as(x0)
        assert len(t) == x0.shape[0], "t has to have batch size dimension"
        eps = self.sample_noise_like(x0)
        xt, ut = self._paired_path_sample(x0, x1, t, eps, i, j)
        self.x0, self.x1 = x0[i], x1[j]
        if return_noise:
            return t, xt, ut, eps
        return t, xt, ut

    def guided_sample_location_and_conditional_flow(self, x0, x1, y0=None, y1=None, t=None, return_noise=False):
        """conditional_flow_matching_dynamic.py:273-315: the labels follow their rows through the OT re-pairing."""
        i, j = self.ot_sampler.sample_map_indices(x0, x1)
        if t is None:
            t = torch.rand(x0.shape[0]).type_as(x0)
        assert len(t) == x0.shape[0], "t has to have batch size dimension"
        eps = self.sample_noise_like(x0)
        xt, ut = self._paired_path_sample(x0, x1, t, eps, i, j)
        self.x0, self.x1 = x0[i], x1[j]
        y0 = y0[i] if y0 is not None else None
        y1 = y1[j] if y1 is not None else None
        if return_noise:
            return t, xt, ut, y0, y1, eps
        return t, xt, ut, y0, y1

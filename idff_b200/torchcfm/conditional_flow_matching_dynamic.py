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

    def sample_location_and_conditional_flow_many(self, x0, x1, return_noise=False):
        """sample_location_and_conditional_flow for S independent minibatches at once: x0, x1 (S, B, *dim) -> t (S, B), xt, ut
        (S, B, *dim) [, eps].  The minibatches of the next S training steps do not depend on the parameters, so their S
        couplings are one batched assignment launch (idff_ot_assign_batch, one CTA per minibatch: 13 us per coupling at B = 256
        instead of 2.6 ms), their index pairs one draw and their path samples one launch over S * B rows with the re-pairing as
        fused gather indices.  Step s of the result is distributed as the reference's per-step call on (x0[s], x1[s])
        (conditional_flow_matching_dynamic.py:268-271)."""
        if x0.dim() < 3 or x0.shape != x1.shape:
            raise ValueError(f"expected x0, x1 of equal shape (S, B, *dim), got {tuple(x0.shape)} and {tuple(x1.shape)}")
        S, B = int(x0.shape[0]), int(x0.shape[1])
        from .optimal_transport import ot_assign_batch   # (CUDA tensors only, like every path sample here: no CPU path)
        perm = ot_assign_batch(x0, x1)
        i = torch.randint(0, B, (S, B), device=x0.device)
        j = torch.gather(perm, 1, i)
        off = torch.arange(S, device=x0.device)[:, None] * B
        fi, fj = (i + off).reshape(-1), (j + off).reshape(-1)
        x0f, x1f = x0.reshape(S * B, *x0.shape[2:]), x1.reshape(S * B, *x1.shape[2:])
        t = torch.rand(S * B, device=x0.device).type_as(x0)
        eps = self.sample_noise_like(x0f)
        xt, ut = self._paired_path_sample(x0f, x1f, t, eps, fi, fj)
        out = (t.view(S, B), xt.view(x0.shape), ut.view(x0.shape))
        return out + (eps.view(x0.shape),) if return_noise else out

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

"""gamma sweeps by MMD -- the caller of the sampler hot path in the reference:

    find_best_gammas_mmd               2D_examples/Autograd_example_order2.py:281-303   (8x8 tuples)
    the 10^3-tuple order-3 sweep       2D_examples/Autograd_example_order3.py:288-314

All tuples of a sweep are solved by ONE sampler launch (idff_sample_rk4_sweep: the reference's 2048 particles fill 32 of
148 SMs per tuple, 64 tuples fill the machine), followed by ONE MMD launch for all tuples (idff_mmd_rbf_sweep; Kxx of the
fixed sample set is evaluated once); the values stay on the device and are read back once at the end (the reference synchronises on `.item()` twice per field call)."""
from __future__ import annotations

import itertools

import torch

from . import fields, mmd, sampler


def _field_for(base_model, g, sigma, impl, score_head=False):
    if score_head:
        if len(g) != 2:
            raise ValueError(f"the score-head field takes (gamma1, gamma2), got {g}")
        model = fields.CustomNetModelScoreHead(base_model, sigma=sigma, gamma1=g[0], gamma2=g[1])
    elif len(g) == 2:
        model = fields.CustomNetModel(base_model, sigma=sigma, gamma1=g[0], gamma2=g[1])
    elif len(g) == 3:
        model = fields.CustomNetModelOrder3(base_model, sigma=sigma, gamma1=g[0], gamma2=g[1], gamma3=g[2])
    else:
        raise ValueError(f"gamma tuple must have 2 or 3 entries, got {g}")
    model.impl = impl
    return model


def sweep_mmd(base_model, x0, true_samples, gamma_tuples, nfe, sigma=0.2, sigma_kernel=1.0, impl=0, chunk=1024, t_span=None,
              score_head=False):
    """MMD(true_samples, odeint(field(gammas), x0, linspace(0,1,nfe+1))[-1]) for every tuple -> list of floats.
    The solves of `chunk` tuples at a time go out as ONE sampler launch (idff_sample_rk4_sweep).  t_span overrides the grid.
    score_head: the order-1 script's field on its MLP(dim=4) (example_order1_with_prediction_head.py:64-96, sweep :305-328)."""
    if t_span is None:
        t_span = torch.linspace(0, 1, nfe + 1)
    true_samples = torch.as_tensor(true_samples, dtype=torch.float32, device=x0.device)
    gamma_tuples = list(gamma_tuples)
    vals = []
    for lo in range(0, len(gamma_tuples), chunk):
        models = [_field_for(base_model, g, sigma, impl, score_head) for g in gamma_tuples[lo:lo + chunk]]
        outs = sampler.sample_rk4_sweep(models, x0, t_span)            # one sampler launch ...
        vals.append(mmd.mmd_sweep(true_samples, outs, sigma_kernel))   # ... and one MMD launch per chunk of tuples
    return torch.cat(vals).cpu().tolist()                              # the only synchronisation


def find_best_gammas_mmd(base_model, x0, true_samples, gamma_range=(0.0, 0.5, 1.0, 2.0), nfe=20, sigma=0.2,
                         sigma_kernel=1.0, order=2, verbose=False, score_head=False):
    """Same contract as the reference: -> (best_gamma_tuple, [((g1, g2[, g3]), mmd), ...]) in product order.  score_head=True:
    the order-1 script's sweep (example_order1_with_prediction_head.py:305-328) over its score-head field."""
    tuples = list(itertools.product(gamma_range, repeat=2 if score_head else order))
    vals = sweep_mmd(base_model, x0, true_samples, tuples, nfe, sigma, sigma_kernel, score_head=score_head)
    results = list(zip(tuples, vals))
    best = min(results, key=lambda r: r[1])
    if verbose:
        for g, v in results:
            print(f"gammas={g} -> MMD: {v:.5f}")
        print(f"best: {best[0]} with MMD={best[1]:.5f}")
    return best[0], results


def compute_mmd_for_gamma_configs(test_model, gamma_configs, x0, t_span, true_samples_tensor, sigma_model=0.2, sigma_kernel=1.0,
                                  device=None, verbose=False):
    """Same contract as the reference (2D_examples/Autograd_example_order3.py:376-425): for each (gamma1, gamma2, gamma3) wrap
    `test_model` in the order-3 field, integrate x0 over t_span with rk4 and return [((g1, g2, g3), mmd), ...] -- here as one
    batched sampler call and one MMD launch per 1024 configurations instead of one solve per configuration."""
    del device
    cfgs = [tuple(g) for g in gamma_configs]
    vals = sweep_mmd(test_model, x0, true_samples_tensor, cfgs, nfe=None, sigma=sigma_model, sigma_kernel=sigma_kernel,
                     t_span=t_span)
    results = list(zip(cfgs, vals))
    if verbose:
        for (g1, g2, g3), v in results:
            print(f"gamma1={g1}, gamma2={g2}, gamma3={g3} -> MMD: {v:.5f}")
    return results

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the IDFF hot path.

A plain torch-CPU restatement (the reference's own language) of every function on the
hot path of MrRezaeiUofT/IDFF, each citing the reference file:line it follows.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`idff_b200/`) never does.

Pinning status ("parity pinned by execution"): the reference ships no tests, seeds or
golden vectors, so this file is pinned against the reference's OWN code
objects executed in the build container (oracle/ref_loader.py + oracle/make_golden.py;
tests/test_oracle_vs_reference.py re-checks bit-exactness whenever /root/reference is
mounted) and against the committed outputs of those runs under tests/golden/.
Two third-party integrators are NOT in /root/reference and not installable here:
  * torchdiffeq (requirements.txt:11, unpinned) fixed-grid `rk4` == 3/8-rule
    `rk4_alt_step_func`  -> `odeint_rk4` below, restated from its published algorithm;
  * torchsde (imported MD_simulation.py:8, not in requirements.txt) `srk` for
    ito/diagonal (Roessler SRI2 tableau)      -> `sdeint_srk` below, Brownian
    increments injected (the reference's own BrownianInterval is entropy-seeded).
  * POT (requirements.txt:10, unpinned) `pot.emd` behind OTPlanSampler.get_map (optimal_transport.py:45,59-94)
    -> `ot_exact_permutation` below: for the uniform equal-size marginals the reference uses, the LP optimum is the assignment
    optimum, obtained with scipy's exact solver.
For those three the parity is "unpinned" (anchored on the reference's call sites only).  The torch optimizer / clip calls of
the training loop (`train_steps_2d`) and numpy's corrcoef / nanmean of `md_compute_metrics` ARE the reference's own calls.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805


# --------------------------------------------------------------------------------------
# a1 / a2 : the networks (torchcfm/models/models.py:4-21, :23-42)
# --------------------------------------------------------------------------------------
def mlp_forward(layers, inp):
    """Linear->SELU x3 ->Linear (models.py:9-21).  layers = [(W[out,in], b[out]) x4] tensors."""
    h = inp
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        h = torch.nn.functional.linear(h, W, b)
        if i + 1 < n:
            h = torch.nn.functional.selu(h)
    return h


def mlp_embedding_forward(layers, emb, inp, ti):
    """MLP_Embedding.forward (models.py:40-42): cat[x, Embedding[ti.long()].squeeze()] -> net."""
    e = torch.nn.functional.embedding(ti.to(torch.long), emb)
    return mlp_forward(layers, torch.cat([inp, e.squeeze()], dim=-1))


def as_layers(weights, dtype=torch.float32):
    """weights: dict with W0..W3,b0..b3 (numpy) -> list of tensor pairs."""
    return [(torch.as_tensor(weights[f"W{i}"]).to(dtype), torch.as_tensor(weights[f"b{i}"]).to(dtype))
            for i in range(4)]


# --------------------------------------------------------------------------------------
# a3 / a4 / a5 / a6 / a7 : the IDFF field wrappers
# --------------------------------------------------------------------------------------
class Field2D(torch.nn.Module):
    """CustomNetModel.forward, order 2 (2D_examples/Autograd_example_order2.py:36-76) and
    order 3 (Autograd_example_order3.py:34-77): nested autograd.grad with ones grad_outputs."""

    def __init__(self, layers, sigma=0.2, gammas=(0.5, 0.5), order=2):
        super().__init__()
        self.layers, self.sigma, self.gammas, self.order = layers, sigma, tuple(gammas), order
        assert len(self.gammas) == order

    def forward(self, t, xt):
        xt = xt.detach().requires_grad_(True)
        sigma_t2 = (self.sigma * torch.sqrt(t * (1 - t))) ** 2                      # order2.py:47
        t_input = t * torch.ones((xt.shape[0], 1), dtype=xt.dtype)                  # :52
        pred = mlp_forward(self.layers, torch.cat([xt, t_input], dim=-1))           # :53-54
        if t == 0:                                                                  # :56-58
            return (pred - xt).detach()
        elif t == 1:                                                                # :59-60
            return ((pred - xt) / 1e-2).detach()
        grads = []
        cur = pred
        for _ in range(self.order):                                                 # :63-70 (order3 :60-71)
            cur = torch.autograd.grad(cur, xt, grad_outputs=torch.ones_like(cur), create_graph=True)[0]
            grads.append(cur)
        G = sum(self.gammas)
        g1 = self.gammas[0]
        out = (torch.sqrt(1 - G * sigma_t2) * (pred - xt) / (1 - t)                 # :73
               + 0.5 * (2 * g1 - 1) * grads[0] * sigma_t2 / (self.sigma ** 2))      # :74
        for k in range(1, self.order):
            out = out + self.gammas[k] * grads[k] * sigma_t2 / (self.sigma ** 2)    # :75 (order3 :75-76)
        return out.detach()


class FieldScoreHead(torch.nn.Module):
    """CustomNetModel.forward of example_order1_with_prediction_head.py:64-96: MLP(dim=4) fed
    cat[xt, xt, t]; out[:, :2] = x1-hat, out[:, 2:] = score head; no autograd, no /sigma^2."""

    def __init__(self, layers, sigma=0.2, gammas=(0.5, 0.5)):
        super().__init__()
        self.layers, self.sigma, self.gammas = layers, sigma, tuple(gammas)

    def forward(self, t, xt):
        sigma_t2 = (self.sigma * torch.sqrt(t * (1 - t))) ** 2                      # :75
        t_input = t * torch.ones((xt.shape[0], 1), dtype=xt.dtype)
        pred = mlp_forward(self.layers, torch.cat([xt, xt, t_input], dim=-1))       # :80-82
        if t == 0:
            return pred[:, :2] - xt                                                 # :86
        elif t == 1:
            return (pred[:, :2] - xt) / 1e-2                                        # :88
        g1, g2 = self.gammas
        return (torch.sqrt(1 - (g1 + g2) * sigma_t2) * (pred[:, :2] - xt) / (1 - t)  # :93
                + 0.5 * (2 * g1 - 1) * pred[:, 2:] * sigma_t2                        # :94
                + g2 * torch.ones_like(xt) * sigma_t2)                               # :95


class FieldCFM(torch.nn.Module):
    """CFMModel.forward (Autograd_example_order2.py:310-332): x1-hat(t, x) - x0 captured at t == 0.  score_head: the order-1
    script's variant on its MLP(dim=4) (example_order1_with_prediction_head.py:335-356): input cat[xt, xt, t], predictions[:, :2]."""

    def __init__(self, layers, score_head=False):
        super().__init__()
        self.layers, self.x0, self.score_head = layers, None, score_head

    def forward(self, t, xt):
        t_input = t * torch.ones((xt.shape[0], 1), dtype=xt.dtype)
        if self.score_head:
            pred = mlp_forward(self.layers, torch.cat([xt, xt, t_input], dim=-1))[:, :xt.shape[1]]
        else:
            pred = mlp_forward(self.layers, torch.cat([xt, t_input], dim=-1))
        if t == 0:
            self.x0 = xt
        return pred - self.x0


class FieldMD(torch.nn.Module):
    """SDE_cal (timeseris_example/MD_simulation.py:481-529): drift f (signs negative, no Gamma in
    the sqrt, t==1 divides by dt) and diagonal diffusion g = 180 sigma^2 sqrt(t(1-t))."""
    noise_type = "diagonal"
    sde_type = "ito"

    def __init__(self, layers, emb, input_dim, sigma=1.0, init_ind=1, dt=0.1, gamma1=1.0, gamma2=1.0):
        super().__init__()
        self.layers, self.emb, self.input_dim = layers, emb, input_dim
        self.sigma, self.dt, self.gamma1, self.gamma2 = sigma, dt, gamma1, gamma2
        self.init_ind = torch.tensor([init_ind])

    def f(self, t, y):
        y = y.view(-1, self.input_dim).detach().requires_grad_(True)                # :498-499
        t_tensor = t * torch.ones((y.shape[0], 1), dtype=y.dtype)                   # :500
        init_tensor = self.init_ind * torch.ones((y.shape[0], 1), dtype=y.dtype)    # :501
        sigma_t2 = (self.sigma ** 2) * t * (1 - t)                                  # :502
        pred = mlp_embedding_forward(self.layers, self.emb, torch.cat([y, t_tensor], dim=-1), init_tensor)
        if t == 0:
            return (pred - y).detach()                                              # :508
        elif t == 1:
            return ((pred - y) / self.dt).detach()                                  # :510
        g1 = torch.autograd.grad(pred, y, grad_outputs=torch.ones_like(pred), create_graph=True)[0]
        g2 = torch.autograd.grad(g1, y, grad_outputs=torch.ones_like(g1), create_graph=True)[0]
        return (torch.sqrt(1 - sigma_t2) * (pred - y) / (1 - t)                     # :523
                - 0.5 * (2 * self.gamma1 - 1) * g1 * sigma_t2 / (self.sigma ** 2)   # :524
                - self.gamma2 * g2 * sigma_t2 / (self.sigma ** 2)).detach()         # :525

    def g(self, t, y):
        return torch.ones_like(y) * 180 * (self.sigma ** 2) * torch.sqrt(t * (1 - t))  # :528-529


# --------------------------------------------------------------------------------------
# a8 : integrator drivers (third-party, restated)
# --------------------------------------------------------------------------------------
_ONE_THIRD = 1.0 / 3.0
_TWO_THIRDS = 2.0 / 3.0


def rk4_38_stage_times(t0, t1):
    """Stage times exactly as torchdiffeq rk_common.rk4_alt_step_func forms them: a 0-d tensor
    times a Python float."""
    dt = t1 - t0
    return dt, (t0, t0 + dt * _ONE_THIRD, t0 + dt * _TWO_THIRDS, t1)


def odeint_rk4(func, y0, t):
    """torchdiffeq.odeint(func, y0, t, method='rk4') for a fixed grid (call sites
    Autograd_example_order2.py:167,213,291): FixedGridODESolver.integrate with grid == t and
    RK4._step_func -> rk4_alt_step_func (3/8 rule).  rtol/atol are ignored by that solver.
    Returns (len(t), *y0.shape)."""
    sol = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt, (ta, tb, tc, td) = rk4_38_stage_times(t0, t1)
        k1 = func(ta, y)
        k2 = func(tb, y + dt * k1 * _ONE_THIRD)
        k3 = func(tc, y + dt * (k2 - k1 * _ONE_THIRD))
        k4 = func(td, y + dt * (k1 - k2 + k3))
        y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        sol.append(y)
    return torch.stack(sol)


# Roessler SRI2 ("srid2") tableau used by torchsde's `srk` for diagonal noise.
_SRI2 = dict(
    C0=(0.0, 1.0, 0.5, 0.0), C1=(0.0, 0.25, 1.0, 0.25),
    A0=((), (1.0,), (0.25, 0.25), (0.0, 0.0, 0.0)),
    A1=((), (0.25,), (1.0, 0.0), (0.0, 0.0, 0.25)),
    B0=((), (0.0,), (1.0, 0.5), (0.0, 0.0, 0.0)),
    B1=((), (-0.5,), (1.0, 0.0), (2.0, -1.0, 0.5)),
    alpha=(1 / 6, 1 / 6, 2 / 3, 0.0),
    beta1=(-1.0, 4 / 3, 2 / 3, 0.0), beta2=(1.0, -4 / 3, 1 / 3, 0.0),
    beta3=(2.0, -4 / 3, -2 / 3, 0.0), beta4=(-2.0, 5 / 3, -2 / 3, 1.0),
)


def sde_step_grid(ts, dt):
    """Fixed-step grid of torchsde BaseSDESolver.integrate: curr_t = min(curr_t + dt, ts[-1]) evaluated on fp32
    tensors (ts is an fp32 tensor, dt a Python float), so the grid accumulates in fp32."""
    grid = [np.float32(ts[0])]
    end = np.float32(ts[-1])
    step = np.float32(dt)
    while grid[-1] < end:
        grid.append(min(np.float32(grid[-1] + step), end))
    return [float(g) for g in grid]


def _linear_interp(t0, y0, t1, y1, t):
    """torchsde _core/interp.py linear_interp: (t1 - t)/(t1 - t0)*y0 + (t - t0)/(t1 - t0)*y1, in fp32."""
    t0, t1, t = np.float32(t0), np.float32(t1), np.float32(t)
    if t1 == t0:
        return y1
    w0 = float(np.float32(t1 - t) / np.float32(t1 - t0))
    w1 = float(np.float32(t - t0) / np.float32(t1 - t0))
    return w0 * y0 + w1 * y1


def sdeint_srk(sde, y0, ts, dt, I_k, I_k0):
    """torchsde.sdeint(sde, y0, ts=ts, dt=dt) default method for ito/diagonal = 'srk'
    (call sites MD_simulation.py:209,228), restated with injected Brownian increments:
    I_k[s] = W(t_{s+1}) - W(t_s), I_k0[s] = the space-time Levy term U of step s
    (shape (n_steps, *y0.shape)).  f/g are evaluated once per stage (torchsde recomputes
    them for j < s; the values are identical).  Outputs at ts by linear interpolation."""
    T = _SRI2
    grid = sde_step_grid(ts, dt)
    ys = [y0]
    y = y0
    prev_t, prev_y = grid[0], y0
    step = 0
    gi = 0
    for out_t in [float(v) for v in ts[1:]]:
        while grid[gi] < out_t:
            t0 = torch.tensor(grid[gi], dtype=y0.dtype)
            t1 = torch.tensor(grid[gi + 1], dtype=y0.dtype)
            h = t1 - t0
            rdt = 1 / h
            sqrt_dt = torch.sqrt(h)
            Ik, Ik0 = I_k[step], I_k0[step]
            Ikk = (Ik ** 2 - h) * 0.5
            Ikkk = (Ik ** 3 - 3 * h * Ik) / 6
            y1 = y
            F, G = [], []
            for s in range(4):
                H0s, H1s = y, y
                for j in range(s):
                    H0s = H0s + T["A0"][s][j] * F[j] * h + T["B0"][s][j] * G[j] * Ik0 * rdt
                    H1s = H1s + T["A1"][s][j] * F[j] * h + T["B1"][s][j] * G[j] * sqrt_dt
                # F[j] / G[j] are f(t0 + C0[j] h, H0[j]) / g(t0 + C1[j] h, H1[j]); stage s needs its
                # own evaluations for later stages and for the update:
                fs = sde.f(t0 + T["C0"][s] * h, H0s)
                gs = sde.g(t0 + T["C1"][s] * h, H1s)
                F.append(fs)
                G.append(gs)
                g_weight = (T["beta1"][s] * Ik + T["beta2"][s] * Ikk / sqrt_dt
                            + T["beta3"][s] * Ik0 * rdt + T["beta4"][s] * Ikkk * rdt)
                y1 = y1 + T["alpha"][s] * fs * h + gs * g_weight
            prev_t, prev_y = grid[gi], y
            y = y1
            gi += 1
            step += 1
        ys.append(y0 if gi == 0 else _linear_interp(prev_t, prev_y, grid[gi], y, out_t))
    return torch.stack(ys)


def sdeint_euler(sde, y0, ts, dt, I_k):
    """Euler-Maruyama on the same grid (torchsde method='euler'), increments injected."""
    grid = sde_step_grid(ts, dt)
    ys = [y0]
    y = y0
    prev_t, prev_y = grid[0], y0
    gi = 0
    for out_t in [float(v) for v in ts[1:]]:
        while grid[gi] < out_t:
            t0 = torch.tensor(grid[gi], dtype=y0.dtype)
            h = torch.tensor(grid[gi + 1], dtype=y0.dtype) - t0
            prev_t, prev_y = grid[gi], y
            y = y + sde.f(t0, y) * h + sde.g(t0, y) * I_k[gi]
            gi += 1
        ys.append(y0 if gi == 0 else _linear_interp(prev_t, prev_y, grid[gi], y, out_t))
    return torch.stack(ys)


# --------------------------------------------------------------------------------------
# a9 / a10 / a11 : conditional path sampling (torchcfm/conditional_flow_matching.py)
# --------------------------------------------------------------------------------------
def pad_t_like_x(t, x):
    """conditional_flow_matching.py:12-33."""
    if isinstance(t, (float, int)):
        return t
    return t.reshape(-1, *([1] * (x.dim() - 1)))


def sqrt_ieee(x):
    """Correctly rounded fp32 square root.  torch's CPU sqrt goes through MKL VML (vsSqrt) for tensors above a
    few hundred elements and is then off by one ulp in ~0.6% of the entries [measured: 6 of 1000]; torch CUDA's
    sqrt, torch CPU on small tensors and the kernels here are IEEE-754 round-to-nearest.  The bit-exact contract
    of the path sample is stated against the IEEE result (= what the reference computes on a GPU)."""
    return torch.sqrt(x.double()).to(x.dtype)


def path_sample(x0, x1, t, eps, sigma, exact_ot, ieee_sqrt=False):
    """sample_xt + compute_conditional_flow (conditional_flow_matching.py:98-148):
    xt = t*x1 + (1-t)*x0 + sigma_t*eps ; ut = x1 - x0.  sigma_t = sigma (:79-96) for the base
    matcher, sigma*sqrt(t(1-t)) (:231-233) for the ExactOT matcher."""
    tp = pad_t_like_x(t, x0)
    mu_t = tp * x1 + (1 - tp) * x0
    sq = sqrt_ieee if ieee_sqrt else torch.sqrt
    sigma_t = sigma * sq(t * (1 - t)) if exact_ot else sigma
    sigma_t = pad_t_like_x(sigma_t, x0)
    return mu_t + sigma_t * eps, x1 - x0


def sample_location_and_conditional_flow(x0, x1, sigma, exact_ot, t=None, return_noise=False,
                                         generator=None):
    """conditional_flow_matching.py:153-193: draw order is t (torch.rand on the CPU generator,
    :184) THEN eps (randn_like, :187)."""
    if t is None:
        t = torch.rand(x0.shape[0], generator=generator).type_as(x0)
    assert len(t) == x0.shape[0], "t has to have batch size dimension"
    eps = torch.randn(x0.shape, dtype=x0.dtype, generator=generator)
    xt, ut = path_sample(x0, x1, t, eps, sigma, exact_ot)
    return (t, xt, ut, eps) if return_noise else (t, xt, ut)


def compute_lambda(t, sigma, exact_ot, dynamic):
    """conditional_flow_matching.py:195-211 (2 sigma_t^2/(sigma^2+1e-8)) vs
    conditional_flow_matching_dynamic.py:198-214 (2 sigma_t/(sigma^2+1e-8))."""
    sigma_t = sigma * torch.sqrt(t * (1 - t)) if exact_ot else sigma
    return (2 * sigma_t / (sigma ** 2 + 1e-8)) if dynamic else (2 * sigma_t ** 2 / (sigma ** 2 + 1e-8))


# --------------------------------------------------------------------------------------
# a12 : losses
# --------------------------------------------------------------------------------------
def flow_loss(vt, x1, t):
    """Autograd_example_order2.py:105 (== order3 :102, order1 :129)."""
    return torch.mean(((1 / (1e-2 + 1 - t[:, None])) * (vt - x1)) ** 2)


def train_steps_2d(layers, x0, x1, t, eps, sigma=0.0, lr=2e-4, clip=1.0, return_grads=False, score_sigma=None):
    """The 2-D training loop, Autograd_example_order2.py:89-112 (= Autograd_example_order3.py:90-109), with the random
    draws passed in: x0, x1, eps [S][B][2], t [S][B].  MLP(dim=2, w=256, time_varying=True) as nn.Sequential of the given
    layers, AdamW(lr) with torch defaults (:80), ExactOT matcher sigma (:81, OT pairing off), clip_grad_norm_(1) (:111).
    score_sigma not None: the order-1 script's loop instead (example_order1_with_prediction_head.py:111-137) on its
    MLP(dim=4): outs = net(cat[xt, xt, t]) (:119), vt | st (:120-121), loss = flow_loss + mean((st - logp)^2) with
    logp = log_prob_xt_given_x1_dimwise(xt, x1, x0, t, score_sigma) (:126-132).
    Returns (per-step losses, per-step total gradient norms, final layers[, unclipped gradients of the last step])."""
    dt = layers[0][0].dtype
    mods = []
    for i, (W, b) in enumerate(layers):
        lin = torch.nn.Linear(W.shape[1], W.shape[0]).to(dt)
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(b)
        mods.append(lin)
        if i + 1 < len(layers):
            mods.append(torch.nn.SELU())
    net = torch.nn.Sequential(*mods)                                              # models.py:9-18
    optimizer = torch.optim.AdamW(net.parameters(), lr=lr)                        # order2.py:80
    losses, norms, grads = [], [], None
    for s in range(t.shape[0]):
        optimizer.zero_grad()                                                     # :91
        xt, ut = path_sample(x0[s], x1[s], t[s], eps[s] if eps is not None else torch.zeros_like(x0[s]), sigma, True,
                             ieee_sqrt=True)                                      # :95
        if score_sigma is None:
            vt = net(torch.cat([xt, t[s][:, None]], dim=-1))                      # :98
            loss = flow_loss(vt, x1[s], t[s])                                     # :105
        else:
            outs = net(torch.cat([xt, xt, t[s][:, None]], dim=-1))                # order1:119
            vt, st = outs[:, :2], outs[:, 2:]                                     # order1:120-121
            grads_lp = log_prob_xt_given_x1_dimwise(xt, x1[s], x0[s], t[s], score_sigma)   # order1:126
            loss = flow_loss(vt, x1[s], t[s]) + torch.mean((st - grads_lp) ** 2)  # order1:129-132
        loss.backward()                                                           # :109
        if return_grads and s == t.shape[0] - 1:
            grads = [p.grad.detach().clone() for p in net.parameters()]
        norms.append(float(torch.nn.utils.clip_grad_norm_(net.parameters(), clip)))   # :111
        optimizer.step()                                                          # :112
        losses.append(float(loss.detach()))
    out_layers = [(mods[2 * i].weight.detach().clone(), mods[2 * i].bias.detach().clone()) for i in range(4)]
    res = (losses, norms, out_layers)
    return res + (grads,) if return_grads else res


def md_train_steps(emb, layers, dataset, idx, t, eps, sigma=0.5, lr=1e-3, return_grads=False, rows=None):
    """TrajectoryGenerator.train_model's loop (timeseris_example/MD_simulation.py:117-144) with the random draws passed in:
    idx [S][B] int64 frame indices (randint(1, T), :128), t [S][B], eps [S][B][D].  MLP_Embedding (models.py:23-42) rebuilt from
    (emb, layers); torch.optim.Adam with defaults (:119); ExactOT matcher sigma (:121, pairing off).  rows = (row0, row1) [S][B]:
    the batch re-paired as the `_dynamic` matcher's sample_plan would (conditional_flow_matching_dynamic.py:268) -- x0 / x1 from those
    dataset rows, the embedding still on index_sel.
    Returns (per-step losses, final (emb, layers)[, gradients of the last step in module.parameters() order])."""
    E = torch.nn.Embedding(emb.shape[0], emb.shape[1]).to(emb.dtype)
    mods = []
    with torch.no_grad():
        E.weight.copy_(emb)
        for i, (W, b) in enumerate(layers):
            lin = torch.nn.Linear(W.shape[1], W.shape[0]).to(W.dtype)
            lin.weight.copy_(W)
            lin.bias.copy_(b)
            mods.append(lin)
            if i + 1 < len(layers):
                mods.append(torch.nn.SELU())
    net = torch.nn.Sequential(*mods)                                              # models.py:30-38
    params = [E.weight] + list(net.parameters())                                  # module.parameters() order
    optimizer = torch.optim.Adam(params, lr=lr)                                   # MD_simulation.py:119
    losses, grads = [], None
    for s in range(idx.shape[0]):
        optimizer.zero_grad()                                                     # :127
        x0, x1 = dataset[idx[s] - 1], dataset[idx[s]]                             # :129-130
        if rows is not None:
            x0, x1 = dataset[rows[0][s]], dataset[rows[1][s]]
        xt, ut = path_sample(x0, x1, t[s], eps[s] if eps is not None else torch.zeros_like(x0), sigma, True, ieee_sqrt=True)   # :132
        x = torch.cat([xt, t[s][:, None]], dim=-1)
        vt = net(torch.cat([x, E(idx[s].to(torch.long)).squeeze()], dim=-1))      # :135, models.py:40-42
        loss = md_periodic_loss(vt, x1)                                           # :137-141
        loss.backward()                                                           # :143
        if return_grads and s == idx.shape[0] - 1:
            grads = [p.grad.detach().clone() for p in params]
        optimizer.step()                                                          # :144
        losses.append(float(loss.detach()))
    out = (E.weight.detach().clone(), [(mods[2 * i].weight.detach().clone(), mods[2 * i].bias.detach().clone()) for i in range(4)])
    return (losses, out, grads) if return_grads else (losses, out)


def log_prob_xt_given_x1_dimwise(x_t, x_1, x_0, t, sigma0_sq):
    """example_order1_with_prediction_head.py:26-51."""
    if t.ndim == 1:
        t = t[:, None]
    sigma0_sq = torch.tensor(sigma0_sq, dtype=x_t.dtype)
    mu_t = t * x_1 + (1 - t) * x_0
    sigma2 = sigma0_sq * t * (1 - t) + 1e-4
    diff = x_t - mu_t
    return -0.5 * (torch.log(2 * torch.pi * sigma2) + (diff ** 2) / sigma2)


def score_loss(st, xt, x1, x0, t, sigma):
    """example_order1_with_prediction_head.py:126,131: mean((st - logp)^2), sigma passed as sigma0_sq."""
    return torch.mean((st - log_prob_xt_given_x1_dimwise(xt, x1, x0, t, sigma)) ** 2)


def md_periodic_loss(vt, x1):
    """MD_simulation.py:137-141."""
    return (torch.mean((vt - x1) ** 2) + torch.mean((vt - (x1 + 360)) ** 2)
            + torch.mean((vt - (x1 - 360)) ** 2))


def ot_exact_permutation(x0, x1):
    """OTPlanSampler.get_map (optimal_transport.py:59-94) for equal batch sizes: pot.emd(unif, unif, cdist^2) -- POT (un-vendored,
    unpinned in requirements.txt:10) solves the transport LP with the network simplex; with uniform equal-size marginals its
    optimum is 1/n times a permutation matrix, obtained here with scipy's assignment solver on the fp64 cost of the fp32
    points.  Returns (perm, total cost): plan[i][perm[i]] = 1/n.  PARITY UNPINNED against POT itself (not installed)."""
    from scipy.optimize import linear_sum_assignment
    a = torch.as_tensor(x0).reshape(len(x0), -1).double()
    b = torch.as_tensor(x1).reshape(len(x1), -1).double()
    M = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1).numpy()
    r, c = linear_sum_assignment(M)
    perm = np.empty(len(r), dtype=np.int64)
    perm[r] = c
    return perm, float(M[r, c].sum())


def md_compute_metrics(y_hat, dataset):
    """TrajectoryGenerator.compute_metrics (MD_simulation.py:275-295): per (trajectory, dimension) Pearson CC, RMSE and MAE
    of the generated series against the data over the T frames, then mean / std (ddof = 0) over the K*D series.
    y_hat (T, K, D), dataset (T, D) numpy."""
    y_hat, dataset = np.asarray(y_hat), np.asarray(dataset)
    cc, rmse, mae = [], [], []
    for ii in range(y_hat.shape[1]):
        for jj in range(y_hat.shape[2]):
            cc.append(np.corrcoef(y_hat[:, ii, jj], dataset[:, jj])[0, 1])                  # :282
            rmse.append(np.sqrt(np.nanmean((y_hat[:, ii, jj] - dataset[:, jj]) ** 2)))      # :284
            mae.append(np.nanmean(np.abs(y_hat[:, ii, jj] - dataset[:, jj])))               # :286
    return {"CC_mean": np.mean(cc), "CC_std": np.std(cc), "RMSE_mean": np.mean(rmse), "RMSE_std": np.std(rmse),
            "MAE_mean": np.mean(mae), "MAE_std": np.std(mae)}


# --------------------------------------------------------------------------------------
# a13 : RBF-kernel MMD
# --------------------------------------------------------------------------------------
def mmd_rbf(x, y, sigma=1.0):
    """compute_mmd_rbf (Autograd_example_order2.py:260-278) == compute_mmd (MD_simulation.py:50-86):
    biased V-statistic with K = exp(-cdist^2 / (2 sigma^2))."""
    def k(a, b):
        return torch.exp(-(torch.cdist(a, b, p=2) ** 2) / (2 * sigma ** 2))
    return (k(x, x).mean() + k(y, y).mean() - 2 * k(x, y).mean()).item()


def mmd_rbf_fp64_direct(x, y, sigma=1.0):
    """fp64 evaluation with the direct (a-b)^2 distance: the accuracy yardstick for K3."""
    x = torch.as_tensor(x).double()
    y = torch.as_tensor(y).double()

    def k(a, b):
        d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
        return torch.exp(-d2 / (2 * sigma ** 2))
    return (k(x, x).mean() + k(y, y).mean() - 2 * k(x, y).mean()).item()


def mmd_partial_sums_fp64(x, y, sigma=1.0, chunk=2048):
    """(sum Kxx, sum Kyy, sum Kxy) in fp64, chunked so large sets stay in memory."""
    x = torch.as_tensor(x).double()
    y = torch.as_tensor(y).double()

    def s(a, b):
        tot = 0.0
        for i in range(0, a.shape[0], chunk):
            d2 = torch.cdist(a[i:i + chunk], b, p=2, compute_mode="donot_use_mm_for_euclid_dist") ** 2
            tot += torch.exp(-d2 / (2 * sigma ** 2)).sum().item()
        return tot
    return s(x, x), s(y, y), s(x, y)


# --------------------------------------------------------------------------------------
# a14 : toy inputs (torchcfm/utils.py)
# --------------------------------------------------------------------------------------
def checkerboard(n):
    """utils.py:74-78 (global numpy RNG) + sample_checkerboard :94-95 (.float())."""
    x1 = np.random.rand(n) * 4 - 2
    x2_ = np.random.rand(n) - np.random.randint(0, 2, n) * 2
    x2 = x2_ + (np.floor(x1) % 2)
    return torch.tensor(np.concatenate([x1[:, None], x2[:, None]], 1) * 2).float()


def eight_gaussians(n, scale=5, var=0.1):
    """utils.py:12-33 + sample_8gaussians :85-86: covariance argument is sqrt(var)*I."""
    m = torch.distributions.multivariate_normal.MultivariateNormal(
        torch.zeros(2), math.sqrt(var) * torch.eye(2))
    s = 1.0 / np.sqrt(2)
    centers = torch.tensor([(1, 0), (-1, 0), (0, 1), (0, -1), (s, s), (s, -s), (-s, s), (-s, -s)]) * scale
    noise = m.sample((n,))
    multi = torch.multinomial(torch.ones(8), n, replacement=True)
    return (centers[multi] + noise).float()

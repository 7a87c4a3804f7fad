"""PolyALA autoregressive generation (TrajectoryGenerator.generate_trajectories / generate_onestep,
timeseris_example/MD_simulation.py:197-230): T sequential SDE solves, each starting from the previous frame, each
with its own embedding row (init_ind = t_idx) and gamma2/(1+t_idx).  Every frame is ONE launch of the fused SDE
sampler over all trajectories; frames never leave the device (the reference round-trips through numpy every frame).
`GraphedGenerator` captures the whole T-frame chain (noise draws included) once as a CUDA graph: at the reference's 5
trajectories the chain is T launch-bound solves, and a replay is one `cudaGraphLaunch`."""
from __future__ import annotations

import torch

from . import fields, sampler


def generate_trajectories(model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1=1.0, gamma2=10.0, dt=0.3,
                          device="cuda", x_init=None, generator=None, method="srk", noise=None):
    """-> y_hat (max_length, num_trajectories, data_dim) on `device`.
    noise: optional callable (t_idx, n_steps) -> (I_k, I_k0) to inject Brownian increments (parity tests)."""
    device = torch.device(device)
    y_hat = torch.empty((max_length, num_trajectories, data_dim), device=device, dtype=torch.float32)
    ts = torch.linspace(0, 1, int(nfe))
    n_steps = sampler.sde_num_steps(ts, dt)
    for t_idx in range(max_length):
        sde = fields.SDE_cal(model, input_dim=data_dim, gamma1=gamma1, gamma2=gamma2 / (1 + t_idx), init_ind=t_idx,
                             max_length=max_length, dt=dt, sigma=sigma, device=device)
        if t_idx == 0:
            x1 = x_init if x_init is not None else torch.randn((num_trajectories, data_dim), device=device,
                                                               generator=generator)
        else:
            x1 = y_hat[t_idx - 1]
        bm = noise(t_idx, n_steps) if noise is not None else None
        traj = sampler.sdeint(sde, x1, ts, dt=dt, method=method, bm=bm, generator=generator)
        y_hat[t_idx] = traj[-1]
    return y_hat


class GraphedGenerator:
    """generate_trajectories captured as one CUDA graph.

    gen = GraphedGenerator(model, K, D, T, nfe, sigma, gamma1, gamma2, dt); y_hat = gen.generate(x_init)
    Every replay draws fresh Brownian increments on the device (graph-safe Philox offsets) unless `noise` tensors are
    given at construction: noise = (I_k, I_k0), each (T, n_steps, K, D), read by the captured solves."""

    def __init__(self, model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1=1.0, gamma2=10.0, dt=0.3,
                 device="cuda", method="srk", noise=None, warmup=1):
        self.device = torch.device(device)
        self.x_init = torch.zeros((num_trajectories, data_dim), device=self.device, dtype=torch.float32)
        self._noise = noise
        kw = dict(device=self.device, x_init=self.x_init, method=method,
                  noise=(lambda t_idx, n: (noise[0][t_idx], noise[1][t_idx])) if noise is not None else None)
        args = (model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1, gamma2, dt)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # allocations, stash buffers and first-launch checks happen here
                generate_trajectories(*args, **kw)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.y_hat = generate_trajectories(*args, **kw)

    def generate(self, x_init=None) -> torch.Tensor:
        """-> y_hat (T, K, D); the tensor is overwritten by the next call."""
        if x_init is not None:
            self.x_init.copy_(x_init)
        self.graph.replay()
        return self.y_hat


def generate_onestep(model, num_trajectories, data_dim, max_length, nfe, sigma, x_prev, gamma1=1.0, gamma2=10.0,
                     generator=None, method="srk"):
    """generate_onestep (MD_simulation.py:215-230): init_ind = 0, dt = 1/nfe, start from x_prev (K, D); returns the
    whole interpolated trajectory (nfe, K, D)."""
    sde = fields.SDE_cal(model, input_dim=data_dim, gamma1=gamma1, gamma2=gamma2, init_ind=0, max_length=max_length,
                         dt=1.0 / nfe, sigma=sigma, device=x_prev.device)
    return sampler.sdeint(sde, x_prev, torch.linspace(0, 1, int(nfe)), dt=1.0 / nfe, method=method, generator=generator)


def compute_metrics(y_hat, dataset):
    """CC / RMSE / MAE between the trajectory mean and the data (MD_simulation.py:275-295), on the device."""
    mean = y_hat.mean(dim=1)
    data = torch.as_tensor(dataset, dtype=torch.float32, device=y_hat.device)
    err = mean - data
    rmse = torch.sqrt((err ** 2).mean())
    mae = err.abs().mean()
    a, b = mean.flatten() - mean.mean(), data.flatten() - data.mean()
    cc = (a * b).sum() / torch.sqrt((a * a).sum() * (b * b).sum())
    return {"CC": float(cc), "RMSE": float(rmse), "MAE": float(mae)}

"""PolyALA autoregressive generation (TrajectoryGenerator.generate_trajectories / generate_onestep,
timeseris_example/MD_simulation.py:197-230): T sequential SDE solves, each starting from the previous frame, each
with its own embedding row (init_ind = t_idx) and gamma2/(1+t_idx).

`generate_trajectories` runs the whole chain in ONE launch of the chain kernel (`idff_md_generate`: one CTA per 4
trajectories, network resident in shared memory) when it serves the shape -- the latency regime of the reference, which
generates 5 trajectories -- and otherwise one launch of the fused SDE sampler per frame over all trajectories; frames never
leave the device (the reference round-trips through numpy every frame).  `GraphedGenerator` captures the per-frame path
(noise draws included) once as a CUDA graph."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, fields, sampler


def chain_serves(model, num_trajectories, data_dim, device) -> bool:
    """Is the one-launch chain kernel the better path for this generation?  It serves hidden width 128, D <= 63; 592
    trajectories run concurrently (15 ms per 200 frames), more in waves -- up to ~3500 trajectories that beats one
    throughput-kernel launch per frame (0.5 ms per frame whatever the count, up to 4096)."""
    pw = fields.SDE_cal(model, input_dim=data_dim, device=device)._packed(torch.device(device))
    return pw.dims[1] == 128 and data_dim <= 63 and pw.dims[0] - pw.emb_dim == data_dim + 1 and 1 <= num_trajectories <= 3552


def generate_trajectories_chain(model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1=1.0, gamma2=10.0, dt=0.3,
                                device="cuda", x_init=None, generator=None, method="srk", noise=None):
    """generate_trajectories (MD_simulation.py:197-213) as ONE kernel launch (idff_md_generate).
    noise: optional (I_k, I_k0) tensors of shape (max_length, n_steps, K, D); drawn on the device otherwise."""
    device = torch.device(device)
    K, D, T = int(num_trajectories), int(data_dim), int(max_length)
    ts = torch.linspace(0, 1, int(nfe)).numpy().astype(np.float32)
    n_steps = sampler.sde_num_steps(ts, dt)
    if noise is None:
        I_k, I_k0 = sampler.brownian_increments(n_steps, ts, dt, (T, K, D), device, generator)   # (n_steps, T, K, D)
        I_k, I_k0 = I_k.transpose(0, 1).contiguous(), I_k0.transpose(0, 1).contiguous()
    else:
        I_k, I_k0 = noise
        I_k = I_k.contiguous()
        I_k0 = None if I_k0 is None else I_k0.contiguous()
    if tuple(I_k.shape) != (T, n_steps, K, D):
        raise _lib.IdffError(f"I_k must be {(T, n_steps, K, D)}, got {tuple(I_k.shape)}")
    if x_init is None:
        x_init = torch.randn((K, D), device=device, generator=generator)
    _lib.require_cuda(x_init, "x_init")
    x_init = x_init.detach().contiguous()
    sde = fields.SDE_cal(model, input_dim=D, gamma1=gamma1, gamma2=gamma2, init_ind=0, max_length=T, dt=dt, sigma=sigma,
                         device=device)
    pw = sde._packed(device)
    p = sde.field_params(D)
    y_hat = torch.empty((T, K, D), device=device, dtype=torch.float32)
    meth = {"srk": 0, "euler": 1}[method]
    _lib.check(_lib.lib().idff_md_generate(pw.handle, C.byref(p), T, _lib.ptr(x_init), ts.ctypes.data_as(C.c_void_p), len(ts),
                                           float(dt), meth, _lib.ptr(I_k), _lib.ptr(I_k0), _lib.ptr(y_hat), K,
                                           _lib.stream_of(x_init)), "idff_md_generate")
    return y_hat


def generate_trajectories_sweep(model, param_sets, num_trajectories, data_dim, max_length, nfe, dt=0.3, device="cuda",
                                x_init=None, generator=None, method="srk", noise=None):
    """generate_trajectories for every (sigma, gamma1, gamma2) of `param_sets` in ONE launch (idff_md_generate_sweep): the loop body
    of TrajectoryGenerator.evaluate_model (MD_simulation.py:166-169) at one nfe.  A generation of the reference's 5 trajectories
    occupies 2 of the 148 SMs, so up to 74 parameter sets cost what one does.
    -> y_hat (max_length, n_sets, num_trajectories, data_dim); slice [:, i] is bit for bit generate_trajectories_chain under
    param_sets[i] on the same start rows and increments.  x_init: (n_sets, K, D) or None (randn); noise: (I_k, I_k0) of shape
    (max_length, n_steps, n_sets, K, D) or None (drawn on the device)."""
    device = torch.device(device)
    S, K, D, T = len(param_sets), int(num_trajectories), int(data_dim), int(max_length)
    if S == 0:
        raise _lib.IdffError("generate_trajectories_sweep: no parameter sets given")
    ts = torch.linspace(0, 1, int(nfe)).numpy().astype(np.float32)
    n_steps = sampler.sde_num_steps(ts, dt)
    if noise is None:
        I_k, I_k0 = sampler.brownian_increments(n_steps, ts, dt, (T, S * K, D), device, generator)   # (n_steps, T, S K, D)
        I_k, I_k0 = I_k.transpose(0, 1).contiguous(), I_k0.transpose(0, 1).contiguous()
    else:
        I_k, I_k0 = noise
        I_k = I_k.contiguous().view(T, n_steps, S * K, D)
        I_k0 = None if I_k0 is None else I_k0.contiguous().view(T, n_steps, S * K, D)
    if tuple(I_k.shape) != (T, n_steps, S * K, D):
        raise _lib.IdffError(f"I_k must be {(T, n_steps, S, K, D)}, got {tuple(I_k.shape)}")
    if x_init is None:
        x_init = torch.randn((S * K, D), device=device, generator=generator)
    _lib.require_cuda(x_init, "x_init")
    x_init = x_init.detach().contiguous().view(S * K, D)
    sdes = [fields.SDE_cal(model, input_dim=D, gamma1=g1, gamma2=g2, init_ind=0, max_length=T, dt=dt, sigma=sg, device=device)
            for sg, g1, g2 in param_sets]
    pw = sdes[0]._packed(device)
    params = (_lib.FieldParams * S)(*[m.field_params(D) for m in sdes])
    y_hat = torch.empty((T, S * K, D), device=device, dtype=torch.float32)
    meth = {"srk": 0, "euler": 1}[method]
    _lib.check(_lib.lib().idff_md_generate_sweep(pw.handle, C.cast(params, C.c_void_p), S, T, _lib.ptr(x_init),
                                                 ts.ctypes.data_as(C.c_void_p), len(ts), float(dt), meth, _lib.ptr(I_k),
                                                 _lib.ptr(I_k0), _lib.ptr(y_hat), K, _lib.stream_of(x_init)),
               "idff_md_generate_sweep")
    return y_hat.view(T, S, K, D)


def generate_overlay_sets(model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1, gamma2, dt=0.3, device="cuda",
                          generator=None):
    """The three generations plot_overlayed_trajectories compares (MD_simulation.py:303-305) -- (gamma1, gamma2), (gamma1, 0),
    (0, 0) -- in one launch: -> (y_hat_full, y_hat_gamma1_only, y_hat_no_gamma), each (max_length, K, D)."""
    ys = generate_trajectories_sweep(model, [(sigma, gamma1, gamma2), (sigma, gamma1, 0.0), (sigma, 0.0, 0.0)], num_trajectories,
                                     data_dim, max_length, nfe, dt=dt, device=device, generator=generator)
    return ys[:, 0], ys[:, 1], ys[:, 2]


def generate_trajectories(model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1=1.0, gamma2=10.0, dt=0.3,
                          device="cuda", x_init=None, generator=None, method="srk", noise=None, chain=None):
    """-> y_hat (max_length, num_trajectories, data_dim) on `device`.
    noise: optional callable (t_idx, n_steps) -> (I_k, I_k0) to inject Brownian increments (parity tests).
    chain: True / False forces / forbids the one-launch chain kernel; None = use it where it serves the shape."""
    device = torch.device(device)
    if chain is None:
        chain = chain_serves(model, num_trajectories, data_dim, device)
    if chain:
        nz = None
        if noise is not None:
            n_steps = sampler.sde_num_steps(torch.linspace(0, 1, int(nfe)), dt)
            pairs = [noise(t_idx, n_steps) for t_idx in range(max_length)]
            nz = (torch.stack([a for a, _ in pairs]), None if pairs[0][1] is None else torch.stack([b for _, b in pairs]))
        return generate_trajectories_chain(model, num_trajectories, data_dim, max_length, nfe, sigma, gamma1, gamma2, dt,
                                           device, x_init, generator, method, nz)
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
        kw = dict(device=self.device, x_init=self.x_init, method=method, chain=False,
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


def _metric_stats(y_hat, dataset):
    """(..., 6) fp64 device tensor [CC_mean, CC_std, RMSE_mean, RMSE_std, MAE_mean, MAE_std] for y_hat (T, ..., K, D)."""
    y = y_hat.to(torch.float64)
    d = torch.as_tensor(dataset, device=y_hat.device).to(torch.float64)
    d = d.view(d.shape[0], *([1] * (y.dim() - 2)), d.shape[1])
    err = y - d
    rmse = torch.sqrt(torch.nanmean(err * err, dim=0))
    mae = torch.nanmean(err.abs(), dim=0)
    ya, da = y - y.mean(dim=0, keepdim=True), d - d.mean(dim=0, keepdim=True)
    cc = (ya * da).sum(0) / torch.sqrt((ya * ya).sum(0) * (da * da).sum(0))
    flat = lambda v: v.reshape(*v.shape[:-2], -1)          # the K * D series of a set
    cc, rmse, mae = flat(cc), flat(rmse), flat(mae)
    return torch.stack([cc.mean(-1), cc.std(-1, unbiased=False), rmse.mean(-1), rmse.std(-1, unbiased=False), mae.mean(-1),
                        mae.std(-1, unbiased=False)], dim=-1)


_METRIC_KEYS = ("CC_mean", "CC_std", "RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std")


def compute_metrics(y_hat, dataset):
    """TrajectoryGenerator.compute_metrics (MD_simulation.py:275-295) without leaving the device: per (trajectory,
    dimension) Pearson CC, RMSE and MAE of the generated series against the data over the T frames, then mean and std
    (ddof = 0) over the K * D series.  y_hat (T, K, D) CUDA tensor, dataset (T, D).  NaN frames are skipped by RMSE / MAE
    (nanmean) and poison CC, as in the reference.  The old keys CC / RMSE / MAE are the *_mean values."""
    if not y_hat.is_cuda:
        raise RuntimeError("compute_metrics needs a CUDA tensor: idff_b200 has no CPU path")
    out = dict(zip(_METRIC_KEYS, _metric_stats(y_hat, dataset).tolist()))          # one device -> host copy
    out.update(CC=out["CC_mean"], RMSE=out["RMSE_mean"], MAE=out["MAE_mean"])
    return out


def compute_metrics_sets(y_hat_sets, dataset):
    """compute_metrics for every parameter set of a sweep at once: y_hat_sets (T, S, K, D) -> list of S dicts (the reference's six
    keys); one device -> host copy for all sets."""
    if not y_hat_sets.is_cuda:
        raise RuntimeError("compute_metrics_sets needs a CUDA tensor: idff_b200 has no CPU path")
    return [dict(zip(_METRIC_KEYS, row)) for row in _metric_stats(y_hat_sets, dataset).tolist()]


def evaluate_model(model, dataset, gamma1_values, gamma2_values, sigma_values, nfes, num_trajectories=5, dt=0.3,
                   device="cuda", generator=None):
    """TrajectoryGenerator.evaluate_model (MD_simulation.py:159-194) without the plots, the KDE-KL and the CSV: for every
    (nfe, sigma, gamma1, gamma2) generate `num_trajectories` chains (all tuples of one nfe in ONE launch where the chain kernel serves
    the network, generate_trajectories_sweep), compute the reference's metrics and the
    RBF-MMD (sigma_kernel = 1, :178) between all data frames and all generated frames.  -> list of dicts with the reference's
    keys (CC_mean .. MAE_std, nfe, sigma, gamma1, gamma2, MMD), in the reference's loop order."""
    device = torch.device(device)
    data = torch.as_tensor(dataset, dtype=torch.float32, device=device)
    T, D = data.shape
    results = []
    from .mmd import mmd_sweep
    for nfe in nfes:
        sets = [(sigma, gamma1, gamma2) for sigma in sigma_values for gamma1 in gamma1_values for gamma2 in gamma2_values]
        if len(sets) > 1 and chain_serves(model, num_trajectories, D, device):
            # every (sigma, gamma1, gamma2) of this nfe in one launch of the chain kernel, 4096 sets at a time
            ys = torch.cat([generate_trajectories_sweep(model, sets[a:a + 4096], num_trajectories, D, T, nfe, dt=dt, device=device,
                                                        generator=generator) for a in range(0, len(sets), 4096)], dim=1)
        else:
            ys = torch.stack([generate_trajectories(model, num_trajectories, D, T, nfe, sg, gamma1=g1, gamma2=g2, dt=dt,
                                                    device=device, generator=generator) for sg, g1, g2 in sets], dim=1)
        # metrics and MMDs of all sets, then one copy to the host each
        stats = compute_metrics_sets(ys, data)
        mmds = mmd_sweep(data, ys.transpose(0, 1).reshape(len(sets), -1, D), 1.0).tolist()
        for (sigma, gamma1, gamma2), m, mmd in zip(sets, stats, mmds):
            m.update({"nfe": nfe, "sigma": sigma, "gamma1": gamma1, "gamma2": gamma2, "MMD": mmd})
            results.append(m)
    return results


class TrajectoryGenerator:
    """The reference's driver class (timeseris_example/MD_simulation.py:88-295) with the same constructor arguments, attributes
    and method names, minus the plots and the KDE-KL column: train_model -> FusedMDTrainer (the whole loop in persistent-kernel
    launches), generate_trajectories / generate_onestep -> the chain kernel / fused SDE sampler, evaluate_model -> one generation
    launch per nfe for the whole (sigma, gamma1, gamma2) grid.  Everything runs on the CUDA device (`use_cuda` is accepted for
    signature compatibility; there is no CPU path).  The dataset is read from data/{dataset_name}_dataset.p like the reference
    does (:105-115), or passed as `dataset` (T, D)."""

    def __init__(self, dataset_name, sigma, nfe, batch_size, num_trajectories, save_dir, model_name, use_cuda=True, dataset=None,
                 device="cuda"):
        import os
        self.dataset_name = dataset_name
        self.sigma = sigma
        self.batch_size = batch_size
        self.num_trajectories = num_trajectories
        self.save_dir = os.path.join(save_dir, dataset_name)
        self.model_name = model_name
        self.rgn_thr = 30
        self.NFE = nfe
        self.num_steps = 50000
        self.vis_step = 10000
        del use_cuda
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.IdffError("TrajectoryGenerator needs a CUDA device: idff_b200 has no CPU path")
        os.makedirs(self.save_dir, exist_ok=True)
        if dataset is None:
            self._load_dataset()
        else:
            self.dataset = torch.as_tensor(dataset).squeeze().float().to(self.device)
            self.DataDim = int(self.dataset.shape[-1])
            self.max_length = int(self.dataset.shape[0])

    def _load_dataset(self):
        import os
        import pickle
        dataset_path = f"data/{self.dataset_name}_dataset.p"
        if not os.path.exists(dataset_path):
            raise FileNotFoundError(f"Dataset file not found: {dataset_path}")
        with open(dataset_path, "rb") as file:
            data = pickle.load(file)
        orig = data["angles"][1].reshape([data["sim_length"], -1])
        self.dataset = torch.tensor(orig).squeeze().float().to(self.device)
        self.DataDim = orig.shape[-1]
        self.max_length = self.dataset.shape[0]

    def train_model(self, steps_per_launch=2048):
        """:117-157 -- Adam on MLP_Embedding(dim, w=128, time_embed=max_length), the ExactOT matcher's path sample, the periodic
        loss; the best model (lowest step loss, :146-148) is saved to save_dir/model_name.pt at the end of the launch it occurred in."""
        import os
        from .torchcfm.models.models import MLP_Embedding
        from .train import FusedMDTrainer
        model = MLP_Embedding(dim=self.DataDim, w=128, time_embed=self.max_length, time_varying=True).to(self.device)
        tr = FusedMDTrainer(model, self.dataset, batch_size=self.batch_size, sigma=self.sigma)
        best_loss, best_state = float("inf"), None
        best_model_path = os.path.join(self.save_dir, f"{self.model_name}.pt")
        done = 0
        while done < self.num_steps:
            n = min(steps_per_launch, self.num_steps - done)
            low = tr.run(n).min().item()
            done += n
            if low < best_loss:
                best_loss, best_state = low, tr.state_dict()
            if done % self.vis_step < n:
                print(f"Step {done}: best loss {best_loss:.3f}")
        if best_state is not None:
            torch.save(best_state, best_model_path)
        print(f"Training completed. Best loss {best_loss:.3f}; state_dict at {best_model_path}")
        self.trainer = tr
        return model

    def generate_trajectories(self, model, gamma1=1, gamma2=10):
        """:197-213 -> numpy (max_length, num_trajectories, DataDim)"""
        return generate_trajectories(model, self.num_trajectories, self.DataDim, self.max_length, self.NFE, self.sigma,
                                     gamma1=gamma1, gamma2=gamma2, dt=0.3, device=self.device).cpu().numpy()

    def generate_onestep(self, model, nfe, time_index, gamma1=1, gamma2=10):
        """:215-230 -> numpy (nfe, num_trajectories, DataDim)"""
        if time_index == 0:
            x1 = torch.randn((self.num_trajectories, self.DataDim), device=self.device)
        else:
            x1 = self.dataset[time_index - 1].unsqueeze(0).repeat(self.num_trajectories, 1).contiguous()
        return generate_onestep(model, self.num_trajectories, self.DataDim, self.max_length, nfe, self.sigma, x1, gamma1=gamma1,
                                gamma2=gamma2).cpu().numpy()

    def compute_metrics(self, y_hat):
        """:275-295 (the six keys of the reference)"""
        m = compute_metrics(torch.as_tensor(y_hat, dtype=torch.float32, device=self.device), self.dataset)
        return {k: m[k] for k in _METRIC_KEYS}

    def evaluate_model(self, model, gamma1_values, gamma2_values, sigma_values, nfes):
        """:159-194 -> list of result rows (also written to save_dir/evaluation_results.csv like the reference's DataFrame)."""
        import csv
        import os
        rows = evaluate_model(model, self.dataset, gamma1_values, gamma2_values, sigma_values, nfes,
                              num_trajectories=self.num_trajectories, dt=0.3, device=self.device)
        path = os.path.join(self.save_dir, "evaluation_results.csv")
        with open(path, "w", newline="") as fh:
            wr = csv.DictWriter(fh, fieldnames=list(rows[0].keys()) if rows else [])
            wr.writeheader()
            wr.writerows(rows)
        print(f"Evaluation results saved to {path}")
        return rows

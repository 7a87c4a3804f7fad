"""Particle-sharded launcher for one NVSwitch box.

Sampling is embarrassingly parallel: rank r owns the contiguous particle rows [r*N/G, (r+1)*N/G) of a globally
defined x0 (generated in fixed 2^20-row chunks, chunk c seeded with base_seed + c, so every row is identical for any
G) and runs the fused sampler on them with no communication.  The only collectives belong to the MMD evaluation:
one all_gather of a fixed-size strided subsample of each rank's generated rows, and one all_gather of the three fp64
partial kernel sums of the row block each rank evaluates; the sums are added in rank order (deterministic)."""
from __future__ import annotations

import torch
import torch.distributed as dist

CHUNK_ROWS = 1 << 20


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_total: int, rank: int, world_size: int):
    """Contiguous block partition; the first n_total % world_size ranks get one extra row."""
    base, extra = divmod(n_total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def global_x0_rows(begin: int, end: int, dim: int, device, base_seed: int = 1000, scale: float = 1.0 / 1.5,
                   chunk_rows: int = CHUNK_ROWS):
    """Rows [begin, end) of the global x0 = randn(N, dim) * scale (Autograd_example_order2.py:136), chunk-seeded."""
    out = torch.empty((end - begin, dim), device=device, dtype=torch.float32)
    gen = torch.Generator(device=device)
    pos = begin
    while pos < end:
        c = pos // chunk_rows
        c0 = c * chunk_rows
        hi = min(end, c0 + chunk_rows)
        gen.manual_seed(base_seed + c)
        chunk = torch.randn((chunk_rows, dim), device=device, generator=gen)
        out[pos - begin:hi - begin] = chunk[pos - c0:hi - c0] * scale
        pos = hi
    return out


def strided_subsample(rows: torch.Tensor, n_eval: int):
    """n_eval rows at a fixed stride (all rows if there are fewer)."""
    n = rows.shape[0]
    if n <= n_eval:
        return rows
    idx = (torch.arange(n_eval, device=rows.device) * (n // n_eval))
    return rows.index_select(0, idx)


def gather_eval_set(local_eval: torch.Tensor, group=None):
    """all_gather of equal-sized per-rank evaluation rows -> (G*n_eval, D), rank-major."""
    rank, ws = world()
    if ws == 1:
        return local_eval
    parts = [torch.empty_like(local_eval) for _ in range(ws)]
    dist.all_gather(parts, local_eval.contiguous(), group=group)
    return torch.cat(parts, 0)


def sharded_mmd(true_x: torch.Tensor, gen_local: torch.Tensor, n_eval_per_rank: int, sigma: float = 1.0,
                pair_sums_fn=None, group=None):
    """MMD^2 between the replicated true set and the all-gathered evaluation subsample of the generated rows.
    Each rank evaluates its row block of Kxx / Kxy (rows of true_x) and Kyy (rows of the gathered set)."""
    if pair_sums_fn is None:
        from .mmd import mmd_partial_sums as pair_sums_fn
    rank, ws = world()
    y_all = gather_eval_set(strided_subsample(gen_local, n_eval_per_rank), group)
    N, M = true_x.shape[0], y_all.shape[0]
    sums = pair_sums_fn(true_x, y_all, sigma, shard_bounds(N, rank, ws), shard_bounds(M, rank, ws))
    sums = sums.to(torch.float64)
    if ws > 1:
        parts = [torch.empty_like(sums) for _ in range(ws)]
        dist.all_gather(parts, sums, group=group)
        sums = torch.stack(parts, 0).sum(0)      # fixed rank order
    val = sums[0] / (float(N) * N) + sums[1] / (float(M) * M) - 2.0 * sums[2] / (float(N) * M)
    return float(val.item()), y_all


class ShardedSampler:
    """Runs `model` (an idff_b200.fields wrapper) over this rank's shard of N global particles."""

    def __init__(self, model, dim: int, device, base_seed: int = 1000, sample_fn=None):
        self.model, self.dim, self.device, self.base_seed = model, dim, device, base_seed
        if sample_fn is None:
            from .sampler import sample_rk4 as sample_fn
        self.sample_fn = sample_fn

    def shard(self, n_total: int):
        rank, ws = world()
        return shard_bounds(n_total, rank, ws)

    def x0(self, n_total: int):
        b, e = self.shard(n_total)
        return global_x0_rows(b, e, self.dim, self.device, self.base_seed)

    def sample(self, n_total: int, t_span, x0=None):
        if x0 is None:
            x0 = self.x0(n_total)
        return self.sample_fn(self.model, x0, t_span)

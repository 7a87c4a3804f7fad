"""Host logic of the particle-sharded launcher on CPU: world_size-2 gloo processes must reproduce the
single-process result (shards, chunk-seeded x0, gathered MMD)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import torch
import torch.multiprocessing as mp

import idff_oracle as O
from idff_b200 import launcher as LA


def _oracle_pair_sums(x, y, sigma, x_rows, y_rows):
    xb, xe = x_rows
    yb, ye = y_rows
    sx, _, sxy = O.mmd_partial_sums_fp64(x[xb:xe], y, sigma) if xe > xb else (0.0, 0.0, 0.0)
    sxx = O.mmd_partial_sums_fp64(x[xb:xe], x, sigma)[2] if xe > xb else 0.0
    syy = O.mmd_partial_sums_fp64(y[yb:ye], y, sigma)[2] if ye > yb else 0.0
    return torch.tensor([sxx, syy, sxy], dtype=torch.float64)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for ws in (1, 2, 3, 8):
            spans = [LA.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_global_x0_is_shard_invariant():
    whole = LA.global_x0_rows(0, 1000, 2, "cpu", chunk_rows=256)
    for ws in (2, 3, 8):
        parts = [LA.global_x0_rows(*LA.shard_bounds(1000, r, ws), 2, "cpu", chunk_rows=256) for r in range(ws)]
        assert torch.equal(torch.cat(parts), whole)
    assert whole.std().item() == pytest.approx(1 / 1.5, rel=0.1)


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n_total = 3001
        b, e = LA.shard_bounds(n_total, rank, ws)
        gen_local = LA.global_x0_rows(b, e, 2, "cpu", chunk_rows=512) * 2.0 + 0.3     # stand-in for sampler output
        g = torch.Generator().manual_seed(0)
        true_x = torch.randn(700, 2, generator=g)
        val, y_all = LA.sharded_mmd(true_x, gen_local, 400, 1.0, pair_sums_fn=_oracle_pair_sums)
        q.put((rank, val, y_all.shape[0], (b, e)))
    finally:
        torch.distributed.destroy_process_group()


def test_sharded_mmd_two_ranks_matches_single_process():
    ws, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                      # every rank holds the same value
    assert res[0][2] == 800
    # single process: same global rows, same strided subsample per shard
    g = torch.Generator().manual_seed(0)
    true_x = torch.randn(700, 2, generator=g)
    ys = []
    for r in range(ws):
        b, e = LA.shard_bounds(3001, r, ws)
        assert (b, e) == res[r][3]
        ys.append(LA.strided_subsample(LA.global_x0_rows(b, e, 2, "cpu", chunk_rows=512) * 2.0 + 0.3, 400))
    ref = O.mmd_rbf_fp64_direct(true_x, torch.cat(ys))
    assert res[0][1] == pytest.approx(ref, abs=1e-12)

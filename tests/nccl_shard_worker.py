"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun, NCCL).  Every rank solves its shard of the global
particle set with the default sampler and takes part in the all-gathered MMD; rank 0 then solves ALL rows alone and checks
that the concatenated shards are bit-identical to its single-GPU result and that the sharded MMD equals the single-process
evaluation of the same gathered set.  Prints one line `MULTIRANK_OK ...` on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import load_golden
    from idff_b200 import fields, launcher, sampler
    from idff_b200.mmd import compute_mmd_rbf
    from idff_b200.weights import PackedWeights
    rank, local, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pw = PackedWeights.from_npz(load_golden("weights_checkerboard"), device=dev)
    model = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    n_total = (1 << 20) + 12345          # crosses a seed chunk, uneven shards
    ts = torch.linspace(0, 1, 3)
    sh = launcher.ShardedSampler(model, 2, dev, base_seed=1000)
    mine = sh.sample(n_total, ts)
    b, e = sh.shard(n_total)
    assert mine.shape == (e - b, 2)
    g = torch.Generator(device=dev).manual_seed(3)
    true_x = torch.randn(4096, 2, device=dev, generator=g)
    val, y_all = launcher.sharded_mmd(true_x, mine, 2048, 1.0)
    # gather every shard on rank 0 (padded to the largest shard)
    sizes = [launcher.shard_bounds(n_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, 2, device=dev)
    pad[: e - b] = mine
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    vals = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(ws)]
    dist.all_gather(vals, torch.tensor([val], dtype=torch.float64, device=dev))
    ok = True
    if rank == 0:
        whole = sampler.sample_rk4(model, launcher.global_x0_rows(0, n_total, 2, dev, 1000), ts)
        cat = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])
        assert torch.equal(cat, whole), "shards differ from the single-GPU solve of the same global rows"
        assert all(v.item() == vals[0].item() for v in vals), "ranks disagree on the MMD value"
        ys = torch.cat([launcher.strided_subsample(whole[lo:hi], 2048) for lo, hi in sizes])
        assert torch.equal(ys, y_all)
        single = compute_mmd_rbf(true_x, ys, 1.0)
        assert abs(single - val) <= 1e-9 + 1e-6 * abs(single), (single, val)
        print(f"MULTIRANK_OK ws={ws} rows={n_total} mmd={val:.6e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

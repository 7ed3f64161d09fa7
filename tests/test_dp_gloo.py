"""Host-side data-parallel logic on CPU (gloo, world_size 2): sharding of the reference's seeded permutation, global loss
scaling and the flat SUM all-reduce reproduce the single-process full-batch gradient.  Per-rank gradients come
from the CPU oracle here (no GPU in this test); the GPU variant of the same check lives in test_gpu_train.py."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opendpd_b200 import dp


def test_shards_partition_the_reference_batch():
    perm = dp.epoch_permutation(1000, seed=0)
    ref = torch.randperm(1000, generator=torch.Generator().manual_seed(0))
    assert torch.equal(perm, ref)
    for B, world in ((64, 2), (64, 8), (10, 4)):
        for step in (0, 3, 1000 // B):       # last one is the partial batch
            parts = [dp.shard_batch_indices(perm, step, B, r, world) for r in range(world)]
            got = torch.cat([p for p, _ in parts])
            assert torch.equal(got, perm[step * B:(step + 1) * B])
            sizes = [p.numel() for p, _ in parts]
            assert max(sizes) - min(sizes) <= 1
            assert all(n == got.numel() for _, n in parts)


def test_gather_frames_matches_reference_framing():
    stream = torch.arange(40, dtype=torch.float32).view(20, 2)
    fr = dp.gather_frames(stream, torch.tensor([0, 5, 16]), 4)
    assert fr.shape == (3, 4, 2) and torch.equal(fr[1], stream[5:9]) and torch.equal(fr[2], stream[16:20])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    rng = np.random.default_rng(0)
    N, T, B, H = 300, 24, 10, 8
    stream = (0.3 * rng.standard_normal((N, 2))).astype(np.float32)
    target = (0.9 * stream).astype(np.float32)
    params = (0.3 * rng.standard_normal(oracle.n_params("dgru", H))).astype(np.float32)
    perm = dp.epoch_permutation(N - T + 1, seed=3)
    idx, n_global = dp.shard_batch_indices(perm, 1, B, rank, world)
    x = dp.gather_frames(torch.from_numpy(stream), idx, T).numpy()
    y = dp.gather_frames(torch.from_numpy(target), idx, T).numpy()
    count = 2.0 * n_global * T
    r = oracle.run("dgru", x, params, target=y, H=H, dtype=np.float64, loss_count=count)
    buf = torch.from_numpy(np.concatenate([r["gparams"], [r["loss"]]]))
    dp.allreduce_flat_(buf)
    if rank == 0:
        idx_all = perm[B:2 * B]
        xa = dp.gather_frames(torch.from_numpy(stream), idx_all, T).numpy()
        ya = dp.gather_frames(torch.from_numpy(target), idx_all, T).numpy()
        full = oracle.run("dgru", xa, params, target=ya, H=H, dtype=np.float64)
        q.put((float(np.abs(buf[:-1].numpy() - full["gparams"]).max() / np.abs(full["gparams"]).max()),
               float(abs(buf[-1].item() - full["loss"]) / full["loss"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dp_allreduce_reproduces_full_batch_gradient(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gerr, lerr = q.get(timeout=10)
    assert gerr < 1e-12 and lerr < 1e-12

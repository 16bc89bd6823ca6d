"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the shard
arithmetic and the final output gather (the only collective on the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acme_jl_b200.distributed import gather_outputs, shard_range


def test_shard_ranges_cover_batch():
    for batch in (1, 2, 7, 8, 65536, 8192 + 3):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                f, c = shard_range(batch, world, r)
                got.extend(range(f, f + c))
            assert got == list(range(batch))
            counts = [shard_range(batch, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, batch, N, ny, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard_range(batch, world, rank)
        # stand-in for the per-shard kernel output: y[b, n, k] = b + 0.001*n + 100*k
        b = torch.arange(first, first + count, dtype=torch.float64).reshape(-1, 1, 1)
        n = torch.arange(N, dtype=torch.float64).reshape(1, -1, 1)
        k = torch.arange(ny, dtype=torch.float64).reshape(1, 1, -1)
        y_local = (b + 0.001 * n + 100 * k).contiguous()
        y = gather_outputs(y_local, batch)
        bb = torch.arange(batch, dtype=torch.float64).reshape(-1, 1, 1)
        ok = torch.equal(y, bb + 0.001 * n + 100 * k)
        # sample-major shards (N, count, ny) gather into (N, B, ny)
        ys = gather_outputs(y_local.transpose(0, 1).contiguous(), batch, layout="sample")
        ok = ok and tuple(ys.shape) == (N, batch, ny) and torch.equal(ys, (bb + 0.001 * n + 100 * k).transpose(0, 1))
        q.put((rank, bool(ok), tuple(y.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_gather_outputs_gloo_world2(batch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, N, ny = 2, 5, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, N, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (batch, N, ny)

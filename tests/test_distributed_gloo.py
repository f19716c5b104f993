"""CPU, world_size 2, gloo: the N > 1 host-side logic (sharding, estimate all-gather, error-sum all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adafortitran_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        full = torch.view_as_complex(torch.randn(6, 120, 14, 2, generator=g))       # same on every rank
        truth = torch.view_as_complex(torch.randn(6, 120, 14, 2, generator=g))
        lo, hi = D.shard_range(6, rank, world)
        local = full[lo:hi].clone()
        gathered = D.gather_estimates(local)
        err = (local - truth[lo:hi]).to(torch.complex128)
        sums = torch.tensor([float((err.abs() ** 2).sum()), float((truth[lo:hi].to(torch.complex128).abs() ** 2).sum())],
                            dtype=torch.float64)
        D.reduce_error_sums(sums)
        ok = torch.equal(gathered, full)
        f64, t64 = full.to(torch.complex128), truth.to(torch.complex128)
        ref = torch.tensor([float(((f64 - t64).abs() ** 2).sum()), float((t64.abs() ** 2).sum())], dtype=torch.float64)
        q.put((rank, bool(ok), bool(torch.allclose(sums, ref, rtol=1e-9)), D.mse_db_from_sums(sums, full.numel())))
    finally:
        dist.destroy_process_group()


def _worker_ragged(rank, world, port, q):
    """Unequal shards (7 samples over 2 ranks): gather_estimates pads for the collective and trims afterwards."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(9)
        full = torch.view_as_complex(torch.randn(7, 12, 4, 2, generator=g))
        lo, hi = D.shard_range(7, rank, world)
        gathered = D.gather_estimates(full[lo:hi].clone())
        q.put((rank, bool(torch.equal(gathered, full)), tuple(gathered.shape)))
    finally:
        dist.destroy_process_group()


def test_gather_unequal_shards_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ragged, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(r[1] and r[2] == (7, 12, 4) for r in results), results


def test_shard_range_partitions_exactly():
    for batch in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert [r[0] for r in results] == [0, 1]
    assert all(r[1] and r[2] for r in results)
    assert abs(results[0][3] - results[1][3]) < 1e-12
    assert np.isfinite(results[0][3])

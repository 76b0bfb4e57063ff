"""Multi-rank host logic on CPU: world_size-2 (and 3, uneven shards) gloo process groups.

Each rank evaluates ITS shard with the CPU oracle (standing in for its GPU — these tests run
without a device), then the product's gather helpers (datainterpolationsnd.jl_b200/parallel.py)
reassemble the global output, which must equal the unsharded evaluation bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dind_b200 as D
from oracle import oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup():
    rng = np.random.default_rng(42)
    ts = [np.cumsum(0.5 + rng.random(n)) for n in (7, 6, 8)]
    u = rng.random((7, 6, 8, 3))
    pts = [rng.uniform(t[0], t[-1], 1001) for t in ts]          # 1001: uneven split
    grid = [np.sort(rng.uniform(t[0], t[-1], n)) for t, n in zip(ts, (9, 5, 11))]
    return ts, u, pts, grid


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ts, u, pts, grid = _setup()
        kinds = ["linear"] * 3
        # point-sharded eval_unstructured
        local, (lo, hi) = D.shard_t_eval(pts, grid=False)
        assert (lo, hi) == D.shard_range(1001, rank, world)
        out_local = O.evaluate(kinds, ts, u, local, derivative_orders=(0, 1, 0))
        full = D.gather_unstructured(out_local, 1001).numpy()
        ref = O.evaluate(kinds, ts, u, pts, derivative_orders=(0, 1, 0))
        ok_u = np.array_equal(full, ref) and full.shape == (1001, 3)
        # slab-sharded eval_grid (last interpolation axis)
        local_g, (glo, ghi) = D.shard_t_eval(grid, grid=True)
        assert len(local_g[0]) == 9 and len(local_g[1]) == 5 and len(local_g[2]) == ghi - glo
        out_slab = O.evaluate(kinds, ts, u, local_g, grid=True)
        full_g = D.gather_grid(out_slab, 3, 11).numpy()
        ref_g = O.evaluate(kinds, ts, u, grid, grid=True)
        ok_g = np.array_equal(full_g, ref_g) and full_g.shape == (9, 5, 11, 3)
        # scalar output (N_out = 0)
        out_s = O.evaluate(kinds, ts, u[..., 0], local_g, grid=True)
        ok_s = np.array_equal(D.gather_grid(out_s, 3, 11).numpy(), ref_g[..., 0])
        flag = torch.tensor([int(ok_u and ok_g and ok_s)])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            results.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_evaluation_and_gather(world):
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results.get(timeout=5) == 1


def test_gather_single_process_is_identity():
    ts, u, pts, grid = _setup()
    out = O.evaluate(["linear"] * 3, ts, u, pts)
    assert np.array_equal(D.gather_unstructured(out, 1001).numpy(), out)
    outg = O.evaluate(["linear"] * 3, ts, u, grid, grid=True)
    assert np.array_equal(D.gather_grid(outg, 3, 11).numpy(), outg)

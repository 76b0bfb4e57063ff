"""One rank of the multi-GPU parity test (started by tests/test_gpu_multi.py through torch.distributed.run, one
process per GPU, NCCL).  BASELINE configs[4]: the 5-D lookup table u (32^5, 4) Float32 is replicated, the point
list is sharded by contiguous index range (dind_shard_range), every rank evaluates its shard with its own handle
and the outputs are all-gathered by dind_allgather_out (NCCL inside libdind.so).  Checks, on every rank:

  * the rank's shard against the CPU oracle on a subsample (coordinates regenerated from the GLOBAL index);
  * the shards tile a single-handle evaluation of the same global points (subsample, bit for bit);
  * plan-sorted == unsorted, bit for bit;
  * gather == concatenation of the shards: own segment bit-identical, every other segment through an exact
    integer checksum exchanged between the ranks;
  * the one-call form (dind_eval_unstructured_allgather: shard evaluated into its rows of the global array, rows
    exchanged chunk by chunk under the next chunk's kernel) gives the same global array;
  * a slab-sharded eval_grid (3-D trilinear) gathers to the single-GPU grid, bit for bit.

    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_worker.py [--n-total 4000000000] [--no-gather]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-total", type=int, default=80_000_000)
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--report", default=None, help="rank 0 writes a JSON summary here")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import dind_b200 as D
    from oracle import oracle as O
    from tests import helpers as H

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    CELL_SORTED, seed = 16, 43
    rng = np.random.default_rng(5)                      # same stream on every rank
    ts = [np.cumsum(0.5 + rng.random(32)).astype(np.float32) for _ in range(5)]
    shape = (32,) * 5 + (4,)
    u = D.f_empty(shape, np.float32, dev)
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(11)
        u.copy_(torch.rand(shape[::-1], generator=gen, device=dev).permute(*reversed(range(6))))
    flat = u.permute(*reversed(range(6)))               # the contiguous memory behind the column-major view
    dist.broadcast(flat, src=0)                         # u replicated (test plumbing, not the data path)
    u_h = u.cpu().numpy()

    n_total = args.n_total
    lo, hi = D.shard_range(n_total, rank, world)
    n_local = hi - lo
    xs_d = H.counter_points_torch(lo, hi, ts, seed, np.float32, dev, overshoot=0.01)
    itp = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d)), device=local)
    out = D.f_empty((n_local, 4), np.float32, dev)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    D.eval_unstructured_(out, itp)                      # warm-up (AoS repack of u)
    ev[0].record()
    D.eval_unstructured_(out, itp, flags=1)
    ev[1].record()
    torch.cuda.synchronize()
    eval_ms = ev[0].elapsed_time(ev[1])

    # --- oracle on a subsample of THIS shard, coordinates regenerated from the global index
    sel_h = lo + rng.integers(0, n_local, 20_000)       # rng state is identical across ranks; lo differs
    xs_h = H.counter_points_np(sel_h, ts, seed, np.float32, overshoot=0.01)
    sel_d = torch.from_numpy(sel_h - lo).to(dev)
    got = out[sel_d].cpu().numpy()
    ref = O.evaluate(["linear"] * 5, ts, u_h, xs_h)
    err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    assert err <= 1e-5, err
    for d in range(5):
        assert np.array_equal(xs_d[d][sel_d].cpu().numpy(), xs_h[d])
    # --- the shards tile a single-handle evaluation of the same global points (bit for bit)
    single = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_h)), device=local)
    assert np.array_equal(D.eval_unstructured(single).cpu().numpy(), got)
    for d in range(5):
        assert np.array_equal(single.interp_dims[d].idx_eval, O.get_idx("linear", ts[d], xs_h[d], dtype=np.float32))
    single.close()
    # --- plan-sorted == unsorted (when the plan fits: ~70 B/point)
    if n_local <= 300_000_000:
        out_s = D.f_empty((n_local, 4), np.float32, dev)
        D.eval_unstructured_(out_s, itp, flags=CELL_SORTED)
        assert torch.equal(out_s, out)
        del out_s
    # exact checksum of every rank's shard (all ranks'), for the gathered arrays below
    def checksum(t):
        return t.contiguous().view(torch.int32).to(torch.int64).sum(dim=0)
    mine = torch.stack([checksum(out[:, c]) for c in range(4)])
    sums = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(sums, mine)

    def check_gathered(full, what):
        assert tuple(full.shape) == (n_total, 4)
        assert torch.equal(full[lo:hi], out), what
        for r in range(world):
            rlo, rhi = D.shard_range(n_total, r, world)
            seg = torch.stack([checksum(full[rlo:rhi, c]) for c in range(4)])
            assert torch.equal(seg, sums[r]), f"{what}: segment of rank {r} differs"

    # --- evaluation + all-gather as ONE collective call (dind_eval_unstructured_allgather): the shard goes straight into
    #     its rows of the global array, the other ranks' rows arrive while the next chunk is computed
    fused_ms = None
    if not args.no_gather:
        full2 = D.f_empty((n_total, 4), np.float32, dev)
        D.eval_unstructured_allgather_(full2, itp, n_total)      # warm-up: communicator, stream, events
        torch.cuda.synchronize()
        full2.fill_(float("nan"))
        dist.barrier()
        ev[2].record()
        D.eval_unstructured_allgather_(full2, itp, n_total)
        ev[3].record()
        torch.cuda.synchronize()
        fused_ms = ev[2].elapsed_time(ev[3])
        check_gathered(full2, "eval_unstructured_allgather")
        full2.fill_(float("nan"))
        D.eval_unstructured_allgather_(full2, itp, n_total, n_chunks=7)     # chunk boundaries that do not divide the shard
        torch.cuda.synchronize()
        check_gathered(full2, "eval_unstructured_allgather, 7 chunks")
        del full2
        torch.cuda.empty_cache()
    itp.close()
    del xs_d
    torch.cuda.empty_cache()

    # --- all-gather of the outputs through the C ABI (NCCL inside libdind.so)
    gather_ms, busbw = None, None
    if not args.no_gather:
        full = D.gather_unstructured(out, n_total)      # warm-up: communicator set-up, NCCL channel allocation
        torch.cuda.synchronize()
        dist.barrier()
        ev[2].record()
        full = D.gather_unstructured(out, n_total)
        ev[3].record()
        torch.cuda.synchronize()
        gather_ms = ev[2].elapsed_time(ev[3])
        check_gathered(full, "dind_allgather_out")
        recv = (n_total - n_local) * 16
        busbw = recv / (gather_ms * 1e-3) / 1e9
        del full
        torch.cuda.empty_cache()

    # --- slab-sharded eval_grid: 3-D trilinear u 64^3 -> 128 x 128 x (16 * world + 3) grid (uneven slabs)
    g_last = 16 * world + 3
    rng = np.random.default_rng(17)   # fresh stream: the draws above depend on the rank's shard size
    tg = [np.cumsum(0.5 + rng.random(64)).astype(np.float32) for _ in range(3)]
    ug = rng.random((64, 64, 64)).astype(np.float32)
    xg = [np.linspace(t[0], t[-1], n).astype(np.float32) for t, n in zip(tg, (128, 128, g_last))]
    local_x, (glo, ghi) = D.shard_t_eval([torch.from_numpy(x).to(dev) for x in xg], grid=True, rank=rank, world=world)
    ug_d = D.f_empty(ug.shape, np.float32, dev)
    ug_d.copy_(torch.from_numpy(ug).to(dev))
    itg = D.NDInterpolation(ug_d, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(tg, local_x)), device=local)
    slab = D.eval_grid(itg)
    whole = D.NDInterpolation(ug_d, tuple(D.LinearInterpolationDimension(t, t_eval=torch.from_numpy(x).to(dev)) for t, x in zip(tg, xg)), device=local)
    ref_grid = D.eval_grid(whole)
    assert torch.equal(ref_grid[:, :, glo:ghi], slab)
    if not args.no_gather:
        gathered = D.gather_grid(slab, 3, g_last, interp=itg)
        torch.cuda.synchronize()
        bad = (gathered != ref_grid).flatten(0, 1).any(dim=0).nonzero().flatten().tolist()
        assert not bad, f"rank {rank}: gathered grid differs in last-axis planes {bad} (own slab {glo}:{ghi})"
    itg.close()
    whole.close()

    stats = torch.tensor([eval_ms, gather_ms or 0.0, err, fused_ms or 0.0], device=dev, dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        rep = {"world": world, "n_total": n_total, "points_per_gpu": n_local, "eval_ms_max": round(float(stats[0]), 3),
               "eval_Gpoints_per_s": round(n_total / (float(stats[0]) * 1e-3) / 1e9, 3),
               "allgather_ms_max": round(float(stats[1]), 3) if gather_ms else None,
               "allgather_busbw_GBps": round((n_total - n_local) * 16 / (float(stats[1]) * 1e-3) / 1e9, 1) if gather_ms else None,
               "eval_plus_allgather_fused_ms_max": round(float(stats[3]), 3) if fused_ms else None,
               "eval_plus_allgather_fused_Gpoints_per_s": round(n_total / (float(stats[3]) * 1e-3) / 1e9, 3) if fused_ms else None,
               "oracle_subsample_max_rel_err": float(f"{float(stats[2]):.3e}"), "ok": True}
        print("MULTI_GPU_REPORT " + json.dumps(rep), flush=True)
        if args.report:
            os.makedirs(os.path.dirname(os.path.abspath(args.report)), exist_ok=True)
            with open(args.report, "w") as f:
                json.dump(rep, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

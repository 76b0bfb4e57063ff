"""Multi-GPU host logic: one process per GPU (torchrun), `u` and knots replicated, the
evaluation points sharded — contiguous index ranges for eval_unstructured, slabs of the last
interpolation axis for eval_grid.  Evaluation needs no collective; the optional gather of the
outputs is a torch.distributed all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests),
issued after the kernels, never inside them.

The reference has no multi-device path at all; this module is the "(5) sharding"
subsystem of the north star.
"""
from __future__ import annotations

import numpy as np

from ._lib import shard_range


def _dist():
    import torch.distributed as dist
    return dist


def rank_world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_t_eval(t_evals, *, grid, rank=None, world=None):
    """This rank's slice of the per-dimension evaluation points.

    eval_unstructured: every dimension's t_eval is cut to the same contiguous index range;
    eval_grid: only the LAST dimension's t_eval is cut (a slab), the others stay whole.
    Returns (local t_evals, (lo, hi))."""
    if rank is None or world is None:
        rank, world = rank_world()
    t_evals = list(t_evals)
    if grid:
        lo, hi = shard_range(int(t_evals[-1].shape[0]), rank, world)
        t_evals[-1] = t_evals[-1][lo:hi]
    else:
        lo, hi = shard_range(int(t_evals[0].shape[0]), rank, world)
        t_evals = [t[lo:hi] for t in t_evals]
    return t_evals, (lo, hi)


def _to_rows(out_local, n_lead):
    """Column-major (lead..., out...) -> row-major tensor (C, L) with L = prod(lead), zero-copy when possible."""
    import torch
    if isinstance(out_local, np.ndarray):
        out_local = torch.from_numpy(np.asfortranarray(out_local))
        # from_numpy keeps the Fortran strides
    nd = out_local.ndim
    rev = out_local.permute(*reversed(range(nd))) if nd > 1 else out_local
    rev = rev.contiguous()   # no copy for column-major inputs
    lead = 1
    for s in out_local.shape[:n_lead]:
        lead *= int(s)
    return rev.reshape(-1, lead)


def _from_rows(rows, shape):
    """Inverse of _to_rows: (C, L) row-major -> column-major tensor of `shape`."""
    nd = len(shape)
    t = rows.reshape(tuple(reversed(shape)))
    return t.permute(*reversed(range(nd))) if nd > 1 else t


def _gather_rows(rows_local, counts, unit, group):
    """all-gather of (C, count_r * unit) row blocks with uneven counts (padded to the maximum)."""
    import torch
    dist = _dist()
    world = len(counts)
    C = rows_local.shape[0]
    cmax = max(counts)
    padded = torch.zeros((C, cmax * unit), dtype=rows_local.dtype, device=rows_local.device)
    padded[:, : rows_local.shape[1]] = rows_local
    gathered = torch.empty((world * C, cmax * unit), dtype=rows_local.dtype, device=rows_local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)   # rank-major concatenation along dim 0
    gathered = gathered.view(world, C, cmax * unit)
    return torch.cat([gathered[r, :, : counts[r] * unit] for r in range(world)], dim=1)


def gather_unstructured(out_local, n_total, group=None):
    """All-gather the per-rank outputs of a point-sharded eval_unstructured.

    out_local: (n_local, out...) column-major (torch tensor or NumPy array);
    returns the global (n_total, out...) column-major torch tensor on every rank."""
    _, world = rank_world(group)
    counts = [shard_range(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in counts]
    rows = _to_rows(out_local, 1)
    full = rows if world == 1 else _gather_rows(rows, counts, 1, group)
    return _from_rows(full, (n_total,) + tuple(out_local.shape[1:]))


def gather_grid(out_local, n_in, n_last_total, group=None):
    """All-gather the per-rank slabs (along the last interpolation axis) of a slab-sharded eval_grid.

    out_local: (g_1..g_{N-1}, g_N_local, out...) column-major; returns (g_1..g_N_total, out...)."""
    import torch
    _, world = rank_world(group)
    counts = [shard_range(n_last_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in counts]
    shape = tuple(int(s) for s in out_local.shape)
    unit = 1
    for s in shape[: n_in - 1]:
        unit *= s
    C = 1
    for s in shape[n_in:]:
        C *= s
    rows = _to_rows(out_local, n_in)                       # (C, g_N_local * unit): slab rows are contiguous per component
    full = rows if world == 1 else _gather_rows(rows, counts, unit, group)
    return _from_rows(full, shape[: n_in - 1] + (n_last_total,) + shape[n_in:])

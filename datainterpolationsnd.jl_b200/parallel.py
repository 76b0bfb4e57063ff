"""Multi-GPU host logic: one process per GPU (torchrun), `u` and knots replicated, the
evaluation points sharded — contiguous index ranges for eval_unstructured, slabs of the last
interpolation axis for eval_grid.  Evaluation needs no collective; the optional gather of the
outputs is dind_allgather_out of the C ABI (NCCL over NVLink inside libdind.so: one ncclAllGather per
component column, straight into the final buffer) for device arrays, and a torch.distributed (gloo)
broadcast loop for the CPU tests — issued after the kernels, never inside them.

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


# ---- NCCL path: the collective lives behind the C ABI (dind_comm_init / dind_allgather_out) ------------------
_COMM_HANDLES = {}   # (group id, dtype, device) -> _Handle that owns an NCCL communicator over the group


def init_comm(handle, group=None):
    """Collective over `group`: give `handle` (a _Handle) an NCCL communicator (dind_comm_init).  Rank 0 makes
    the unique id with dind_comm_get_unique_id, torch.distributed only carries the 128 bytes to the other ranks."""
    import ctypes as C

    from . import _lib
    dist = _dist()
    rank, world = rank_world(group)
    lib = _lib.load()
    buf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.dind_comm_get_unique_id(buf))
    box = [bytes(buf)]
    if world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
    uid = (C.c_ubyte * _lib.UNIQUE_ID_BYTES).from_buffer_copy(box[0])
    _lib.check(lib.dind_comm_init(handle.h, uid, rank, world))
    return handle


def _comm_handle(group, np_dtype, device):
    """A communicator-only handle for gathers that are not tied to an NDInterpolation."""
    from . import _Handle
    key = (id(group) if group is not None else None, np.dtype(np_dtype).str, int(device))
    h = _COMM_HANDLES.get(key)
    if h is None:
        h = init_comm(_Handle(1, np_dtype, device), group)
        _COMM_HANDLES[key] = h
    return h


def _gather_device(out_local, lead_shape, n_total, unit, group, interp):
    """dind_allgather_out on column-major CUDA tensors: every component column goes straight into the final
    buffer — the global output is resident exactly once per GPU."""
    import ctypes as C

    import torch

    from . import _is_f_contiguous_torch, _lib, f_empty
    if not _is_f_contiguous_torch(out_local):
        raise _lib.DindError(_lib.ERR_INVALID_ARGUMENT, "device outputs must be column-major (see f_empty)")
    np_dtype = np.float32 if out_local.dtype == torch.float32 else np.float64
    dev = out_local.device.index if out_local.device.index is not None else torch.cuda.current_device()
    if interp is not None:
        h = interp._h
        if not getattr(h, "_comm_ready", False):
            init_comm(h, group)
            h._comm_ready = True
    else:
        h = _comm_handle(group, np_dtype, dev)
    tail = tuple(int(s) for s in out_local.shape[len(lead_shape):])
    n_comp = 1
    for s in tail:
        n_comp *= s
    full = f_empty(tuple(lead_shape[:-1]) + (n_total,) + tail, np_dtype, out_local.device)
    # the gather runs on the handle's stream: order it after the producer of out_local on torch's current stream
    cur = torch.cuda.current_stream(out_local.device)
    h.set_stream(cur.cuda_stream)
    _lib.check(h.lib.dind_allgather_out(h.h, C.c_void_p(out_local.data_ptr()), C.c_void_p(full.data_ptr()),
                                        int(n_total), int(unit), int(n_comp), _lib.FLAG_ASYNC))
    return full


def eval_unstructured_allgather_(out_global, interp, n_total, *, derivative_orders=None, n_chunks=0, group=None):
    """Point-sharded eval_unstructured! + all-gather in one collective call (dind_eval_unstructured_allgather): `interp`
    holds this rank's shard of the points (shard_t_eval), the shard is evaluated straight into its rows of the global
    (n_total, out...) column-major CUDA tensor `out_global`, and the other ranks' rows arrive by NCCL while the next
    chunk of the shard is still being computed.  No local output array.  Returns out_global."""
    import ctypes as C

    import torch

    from . import _is_f_contiguous_torch, _lib, _orders
    if not (_is_cuda_tensor(out_global) and _is_f_contiguous_torch(out_global)):
        raise _lib.DindError(_lib.ERR_INVALID_ARGUMENT, "out_global must be a column-major CUDA tensor (see f_empty)")
    want = (int(n_total),) + tuple(interp.output_size)
    if tuple(out_global.shape) != want:
        raise _lib.DindError(_lib.ERR_SIZE_MISMATCH, f"out_global has shape {tuple(out_global.shape)}, expected {want}")
    h = interp._h
    _, world = rank_world(group)
    if world > 1 and not getattr(h, "_comm_ready", False):
        init_comm(h, group)
        h._comm_ready = True
    h.set_stream(torch.cuda.current_stream(out_global.device).cuda_stream)
    _lib.check(h.lib.dind_eval_unstructured_allgather(h.h, C.c_void_p(out_global.data_ptr()), int(n_total),
                                                      _orders(derivative_orders, interp.N_in), int(n_chunks), _lib.FLAG_ASYNC))
    return out_global


def _gather_rows(rows_local, counts, unit, group):
    """gloo / CPU path (tests): rank r broadcasts its (C, count_r * unit) block into a scratch of that size, which
    is copied into the final (C, total * unit) buffer — the global output is held once."""
    import torch
    dist = _dist()
    rank, world = rank_world(group)
    C = rows_local.shape[0]
    total = sum(counts)
    full = torch.empty((C, total * unit), dtype=rows_local.dtype, device=rows_local.device)
    lo = 0
    for r in range(world):
        n = counts[r] * unit
        if n:
            blk = rows_local.contiguous() if r == rank else torch.empty((C, n), dtype=rows_local.dtype, device=rows_local.device)
            src = dist.get_global_rank(group, r) if group is not None else r
            dist.broadcast(blk, src=src, group=group)
            full[:, lo:lo + n] = blk
        lo += n
    return full


def _is_cuda_tensor(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def gather_unstructured(out_local, n_total, group=None, interp=None):
    """All-gather the per-rank outputs of a point-sharded eval_unstructured.

    out_local: (n_local, out...) column-major (torch tensor or NumPy array);
    returns the global (n_total, out...) column-major torch tensor on every rank.
    CUDA tensors go through dind_allgather_out (NCCL inside libdind.so; `interp`'s handle owns the communicator,
    or a cached communicator-only handle); CPU arrays through torch.distributed (gloo)."""
    _, world = rank_world(group)
    if _is_cuda_tensor(out_local):
        return _gather_device(out_local, tuple(out_local.shape[:1]), n_total, 1, group, interp)
    counts = [shard_range(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in counts]
    rows = _to_rows(out_local, 1)
    full = rows if world == 1 else _gather_rows(rows, counts, 1, group)
    return _from_rows(full, (n_total,) + tuple(out_local.shape[1:]))


def gather_grid(out_local, n_in, n_last_total, group=None, interp=None):
    """All-gather the per-rank slabs (along the last interpolation axis) of a slab-sharded eval_grid.

    out_local: (g_1..g_{N-1}, g_N_local, out...) column-major; returns (g_1..g_N_total, out...)."""
    _, world = rank_world(group)
    shape = tuple(int(s) for s in out_local.shape)
    unit = 1
    for s in shape[: n_in - 1]:
        unit *= s
    if _is_cuda_tensor(out_local):
        return _gather_device(out_local, shape[:n_in], n_last_total, unit, group, interp)
    counts = [shard_range(n_last_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in counts]
    rows = _to_rows(out_local, n_in)                       # (C, g_N_local * unit): slab rows are contiguous per component
    full = rows if world == 1 else _gather_rows(rows, counts, unit, group)
    return _from_rows(full, shape[: n_in - 1] + (n_last_total,) + shape[n_in:])

#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 evaluation engine (BASELINE.json metric:
Gpoints/s for eval_unstructured / eval_grid!, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c4b|c5] [--impl reference]
                    [--only-headline]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline (`value`) is
  c2 (BASELINE configs[1]): 3-D trilinear scalar field u = 256^3 Float32, eval_grid! onto a 512^3 t_eval
      grid with cached per-axis indices/weights, a fresh `u` each step.
The same JSON line also carries every other BASELINE config at its named size:
  N = 1:  "configs": {c1, c3, c3_sorted, c3_plan_order, c3_gradient, c4, c4b, c5, c5_sorted, c5_plan_order}
          (c3: 10^8 points; c5: 5*10^8 points = one GPU's share of the 4*10^9-point config on 8 GPUs);
  N > 1:  "point_sharded": {c5, c3} — the point list sharded by contiguous index range over the N ranks (weak
          scaling: 5*10^8 / 10^8 points per GPU, so c5 at N = 8 IS the named 4*10^9-point config), timed without
          and with the NCCL all-gather of the outputs (dind_allgather_out), max over ranks.
Multi-GPU (torchrun, one rank per GPU): grids are sharded by slab along the last axis, point lists by contiguous
index range, u and knots replicated, no data-path collective.  Weak scaling of the grid headline: the global last
axis is N sweeps over the same knot span (a legitimate unsorted t_eval), so every rank evaluates exactly the N = 1
problem — same u extent, same density, same algorithmic bytes.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, OpenMP over all host cores: the reference itself is Julia and cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpoints/s"

def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, parsed from the newest committed
    `ncu --set full` summary profiles/r<round>_<key>.md (scripts/ncu_summary.py wrote it); None without a capture."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{key}.md")), reverse=True)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for fn in files:
        tot, seen = 0.0, 0
        for line in open(fn):
            m = re.match(r"\| dram__bytes_(read|write)\.sum \| ([0-9.,]+) \| (\w+) \|", line)
            if m:
                tot += float(m.group(2).replace(",", "")) * unit.get(m.group(3), 1.0)
                seen += 1
                if seen == 2:   # the first kernel section = the dominant kernel
                    return {"bytes": int(tot), "source": os.path.relpath(fn, ROOT)}
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ---- workloads -------------------------------------------------------------------------------------
def knots_1d(rng, n, dtype):
    return np.cumsum(0.5 + rng.random(n)).astype(dtype)   # docs/src/usage.md:11-12 style, non-uniform


class Workload:
    name = ""
    dtype = np.float32
    kinds = ()
    degrees = None


class Grid(Workload):
    """eval_grid! workloads: slab-sharded along the last interpolation axis."""
    grid = True
    nurbs = False
    n_u_buffers = 4            # u rotates over several buffers so it is never L2-resident at step start

    def __init__(self, scale=1.0):
        if scale != 1.0:
            self.n_grid = tuple(max(8, int(g * scale)) for g in self.n_grid)
            self.n_u = tuple(max(8, int(g * scale)) for g in self.n_u)

    def points_per_step(self):
        return int(np.prod(self.n_grid))

    def alg_bytes(self):
        # out write + one read of u (and of the NURBS weights) per launch
        esz = np.dtype(self.dtype).itemsize
        return self.points_per_step() * esz + int(np.prod(self.n_u)) * esz * (2 if self.nurbs else 1)

    def n_knots(self, d):
        return self.n_u[d] - self.degrees[d] + 1 if self.kinds[d] == "bspline" else self.n_u[d]

    def build(self, rank, world):
        rng = np.random.default_rng(1234)
        nd = len(self.n_u)
        self.ts = [knots_1d(rng, self.n_knots(d), self.dtype) for d in range(nd)]
        # global grid: the last axis has n_grid[-1] * world points = `world` sweeps over the knot span (t_eval need not
        # be sorted), rank r owns slab r = one sweep: every rank evaluates exactly the N = 1 problem
        self.t_evals = [np.linspace(t[0], t[-1], g).astype(self.dtype) for t, g in zip(self.ts, self.n_grid)]

    def make_dims(self, D, t_evals):
        dims = []
        for k, p, t, x in zip(self.kinds, self.degrees, self.ts, t_evals):
            if k == "linear":
                dims.append(D.LinearInterpolationDimension(t, t_eval=x))
            elif k == "constant":
                dims.append(D.ConstantInterpolationDimension(t, left=True, t_eval=x))
            else:
                dims.append(D.BSplineInterpolationDimension(t, p, t_eval=x))
        return tuple(dims)


class C2(Grid):
    """3-D trilinear, u 256^3 F32 -> 512^3 grid (per GPU)."""
    name = "c2: 3-D trilinear u=256^3 Float32, eval_grid! onto 512^3 (cached idx, new u each step)"
    key, dtype, dtype_name = "c2", np.float32, "f32"
    kinds, degrees = ("linear",) * 3, (0, 0, 0)
    n_u = (256, 256, 256)
    n_grid = (512, 512, 512)   # per GPU


class C4(Grid):
    """4-D degree-2 NURBS, u 64^4 + weights 64^4 F64 -> 128^4 grid (per GPU)."""
    name = "c4: 4-D NURBS (degree 2, NURBSWeights) u=64^4 Float64, eval_grid! onto 128^4"
    key, dtype, dtype_name = "c4", np.float64, "f64"
    kinds, degrees = ("bspline",) * 4, (2, 2, 2, 2)
    n_u = (64,) * 4
    n_grid = (128,) * 4
    nurbs = True
    n_u_buffers = 2


class C4b(C4):
    """Variant with mixed kinds (no reference semantics): axis 3 is Constant(left=true)."""
    name = "c4b: 4-D NURBS (NURBSWeights), mixed axes (BSpline2, BSpline2, Constant(left=true), BSpline2) u=64^4 Float64 -> 128^4"
    key = "c4b"
    kinds, degrees = ("bspline", "bspline", "constant", "bspline"), (2, 2, 0, 2)


class Unstructured(Workload):
    grid = False

    def points_per_step(self):
        return self.n_points

    def build(self, rank, world):
        rng = np.random.default_rng(1234)
        self.ts = [knots_1d(rng, n, self.dtype) for n in self.n_knots]
        self.rank, self.world = rank, world


class C1(Unstructured):
    name = "c1: 2-D linear u=(64,64,2) Float64, eval_unstructured at 1e6 zipped points"
    key, dtype, dtype_name = "c1", np.float64, "f64"
    kinds, degrees, n_knots, n_u, n_points = ("linear",) * 2, (0, 0), (64, 64), (64, 64, 2), 10 ** 6

    def alg_bytes(self):
        return self.n_points * 32 + 64 * 64 * 2 * 8


class C3(Unstructured):
    name = "c3: 3-D cubic BSpline u=(128,128,128,3) Float64, eval_unstructured"
    key, dtype, dtype_name = "c3", np.float64, "f64"
    kinds, degrees, n_knots, n_u = ("bspline",) * 3, (3, 3, 3), (126,) * 3, (128, 128, 128, 3)

    def __init__(self, n_points):
        self.n_points = n_points

    def alg_bytes(self):
        return self.n_points * 48 + 128 ** 3 * 3 * 8


class C5(Unstructured):
    name = "c5: 5-D linear lookup table u=(32^5,4) Float32, eval_unstructured"
    key, dtype, dtype_name = "c5", np.float32, "f32"
    kinds, degrees, n_knots, n_u = ("linear",) * 5, (0,) * 5, (32,) * 5, (32,) * 5 + (4,)

    def __init__(self, n_points):
        self.n_points = n_points

    def alg_bytes(self):
        return self.n_points * 36 + 32 ** 5 * 4 * 4


# ---- clocks sampling -----------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.nv:
            self.th.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---- GPU arm ------------------------------------------------------------------------------------------
def run_gpu(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import dind_b200 as D

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    wl.build(rank, world)
    stream = torch.cuda.current_stream()
    tdt = torch.float32 if wl.dtype == np.float32 else torch.float64

    def to_dev(a):
        t = D.f_empty(a.shape, wl.dtype, dev)
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(dev))
        return t

    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)
    if wl.grid:
        dims = wl.make_dims(D, [to_dev(x) for x in wl.t_evals])
        rev = tuple(reversed(range(len(wl.n_u))))
        u_bufs = []
        for _ in range(wl.n_u_buffers):
            ub = D.f_empty(wl.n_u, wl.dtype, dev)
            ub.copy_(torch.rand(wl.n_u[::-1], generator=gen, device=dev, dtype=tdt).permute(*rev))
            u_bufs.append(ub)
        cache = None
        if wl.nurbs:
            wts = D.f_empty(wl.n_u, wl.dtype, dev)
            wts.copy_((torch.rand(wl.n_u[::-1], generator=gen, device=dev, dtype=tdt) + 0.5).permute(*rev))
            cache = D.NURBSWeights(wts)
        itp = D.NDInterpolation(u_bufs[0], dims, cache=cache, device=local_rank, stream=stream.cuda_stream)
        out = D.f_empty(wl.n_grid, wl.dtype, dev)

        def step(i):
            itp.set_u(u_bufs[i % len(u_bufs)])          # "repeated with new u": borrowed device pointer, no copy
            D.eval_grid_(out, itp, flags=1)             # DIND_FLAG_ASYNC: stay on the stream, timed with CUDA events
    else:
        lo, hi = D.shard_range(wl.n_points * world, rank, world)   # weak scaling: n_points per rank
        n = hi - lo
        mk = {"linear": lambda t, p, x: D.LinearInterpolationDimension(t, t_eval=x),
              "bspline": lambda t, p, x: D.BSplineInterpolationDimension(t, p, t_eval=x, max_derivative_order_eval=1)}
        xs = []
        for t in wl.ts:
            x = torch.rand(n, generator=gen, device=dev, dtype=tdt) * float(t[-1] - t[0]) + float(t[0])
            xs.append(x)
        dims = tuple(mk[k](t, p, x) for k, t, p, x in zip(wl.kinds, wl.ts, wl.degrees, xs))
        shape = tuple(wl.n_u)
        ud = D.f_empty(shape, wl.dtype, dev)
        ud.copy_(torch.rand(shape[::-1], generator=gen, device=dev, dtype=tdt).permute(*reversed(range(len(shape)))))
        itp = D.NDInterpolation(ud, dims, device=local_rank, stream=stream.cuda_stream)
        out = D.f_empty((n,) + shape[len(wl.kinds):], wl.dtype, dev)

        uflags = 1 | (16 if args.sorted else 0) | (32 if args.plan_order else 0)   # ASYNC | CELL_SORTED | PLAN_ORDER
        n_in = len(wl.kinds)
        if args.gradient:
            gout = D.f_empty((n,) + shape[n_in:] + (n_in + 1,), wl.dtype, dev)
            unit = [tuple(1 if d + 1 == s else 0 for d in range(n_in)) for s in range(n_in + 1)]

        def step(i):
            if args.gradient == "fused":        # value + all first partials in one pass
                D.eval_unstructured_gradient(itp, gout, flags=uflags)
            elif args.gradient == "separate":   # what the reference needs: N_in + 1 eval_unstructured! calls
                for s_ in range(n_in + 1):
                    D.eval_unstructured_(gout[..., s_], itp, derivative_orders=unit[s_], flags=uflags)
            else:
                D.eval_unstructured_(out, itp, flags=uflags)

    if args.gather and world > 1:
        # optional NCCL all-gather of the outputs after the kernel (north_star (5)): never inside the kernel loop
        inner_step = step
        n_total = (wl.n_grid[-1] if wl.grid else out.shape[0]) * world

        def step(i):
            inner_step(i)
            if wl.grid:
                D.gather_grid(out, len(wl.n_grid), n_total)
            else:
                D.gather_unstructured(out, n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the first call builds the device-side caches (per-axis tables, AoS copy, cell-sorted plan): time it apart
    torch.cuda.synchronize()
    t_first = time.perf_counter()
    step(0)
    torch.cuda.synchronize()
    first_call_ms = (time.perf_counter() - t_first) * 1e3
    for i in range(1, args.warmup):
        step(i)
    barrier()
    # A 0.1 ms step is launch-bound from Python (set_u + eval_grid! are two ctypes calls, ~0.1 ms per step when several
    # ranks share the host): the steps of one rotation over the u buffers are captured once in a CUDA graph — the API
    # calls are capturable (no allocation, no synchronisation with DIND_FLAG_ASYNC and device arrays) — and the timed
    # region replays it.  Steps that do not fill a graph run eagerly; --no-graph times the eager loop.
    graph, per_graph, launch_mode = None, 0, "eager (one set_u + eval call per step from Python)"
    l0 = itp.launch_count
    step(args.warmup)
    launches_per_step = itp.launch_count - l0
    if wl.grid and not args.no_graph and not (args.gather and world > 1):
        try:
            per_graph = len(u_bufs)
            cap_stream = torch.cuda.Stream(device=dev)
            cap_stream.wait_stream(stream)
            itp.set_stream(cap_stream.cuda_stream)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap_stream):
                for i in range(per_graph):
                    step(i)
            graph = g
            launch_mode = f"cuda graph of {per_graph} steps (one per u buffer), replayed; remainder eager"
        except Exception as exc:   # noqa: BLE001 — capture is an optimisation of the measurement, not of the product
            graph, launch_mode = None, f"eager (graph capture failed: {type(exc).__name__})"
        finally:
            itp.set_stream(stream.cuda_stream)
        torch.cuda.synchronize()
        if graph is not None:
            graph.replay()   # warm
            torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        done = 0
        if graph is not None:
            for _ in range(args.steps // per_graph):
                graph.replay()
            done = (args.steps // per_graph) * per_graph
        for i in range(done, args.steps):
            step(args.warmup + i)
        ev1.record(stream)
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    ms_per_step_rank = ms_total / args.steps
    launches = launches_per_step * args.steps
    kernel = itp.last_kernel

    # ---- cache construction as its own timed path: new t_eval (borrowed device pointers) ->
    # everything the evaluation needs that depends on t_eval.  Grids: the per-axis tables; unstructured: the
    # cell-sorted plan (key kernel + radix sort + sorted copies of t_eval + inverse permutation).
    cache_build = None
    if world == 1 and not args.no_cache_build:
        t_cur = [d.t_eval for d in itp.interp_dims]
        cb_flags = 1 | (0 if wl.grid else 16)
        # timed on the host clock around a device synchronisation: cache construction includes host-side work
        # (scratch allocation for the radix sort) that a stream event would not see
        reps, cb_times = 5, []
        for r in range(reps + 1):
            torch.cuda.synchronize()
            _t0 = time.perf_counter()
            itp.set_t_eval(t_cur)
            itp.build_caches(grid=wl.grid, flags=cb_flags)
            torch.cuda.synchronize()
            cb_times.append((time.perf_counter() - _t0) * 1e3)
        cb_ms = float(np.median(cb_times[1:]))
        n_cb = sum(int(t.shape[0]) for t in t_cur) if wl.grid else int(t_cur[0].shape[0])
        cache_build = {"ms": round(cb_ms, 4), "what": "per-axis index/weight tables" if wl.grid else "cell-sorted plan (keys, radix sort, sorted t_eval copies, inverse permutation)",
                       "entries": n_cb, "Gentries_per_s": round(n_cb / (cb_ms * 1e-3) / 1e9, 4),
                       "evaluations_to_amortise": round(cb_ms / ms_per_step_rank, 2) if not wl.grid else None}

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e_ms, h2d, d2h = None, 0, 0
    if wl.grid:
        u_host = [torch.rand(wl.n_u[::-1], dtype=tdt).pin_memory() for _ in range(2)]
        out_host = torch.empty(wl.n_grid[::-1], dtype=tdt).pin_memory()
        u_np = [t.numpy().transpose(*rev) for t in u_host]          # column-major views of the pinned buffers
        out_np = out_host.numpy().transpose(*rev)
        itp_h = D.NDInterpolation(u_np[0], wl.make_dims(D, wl.t_evals),
                                  cache=None if not wl.nurbs else D.NURBSWeights(wts),
                                  device=local_rank, stream=stream.cuda_stream)
        n_e2e = max(2, min(args.steps, 5))
        for i in range(2):
            itp_h.set_u(u_np[i % 2]); D.eval_grid_(out_np, itp_h)
        barrier()
        # every step uploads its own u from pinned host memory and downloads its own result into pinned host
        # memory; the calls are queued with DIND_FLAG_ASYNC so that the D2H of step i (copy stream) overlaps the
        # H2D + kernel of step i + 1, and the host waits once at the end
        out_host2 = torch.empty(wl.n_grid[::-1], dtype=tdt).pin_memory()
        outs_np = [out_np, out_host2.numpy().transpose(*rev)]
        n_e2e = max(4, min(args.steps, 10))
        for i in range(2):                                  # untimed: the pipelined path's one-time set-up (copy stream,
            itp_h.set_u(u_np[i % 2], flags=1)               # events, second 537 MB staging buffer) — with 8 ranks on one
            D.eval_grid_(outs_np[i % 2], itp_h, flags=1)    # host these cost ~0.3 s and were inside the timed region
        itp_h.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            itp_h.set_u(u_np[i % 2], flags=1)               # H2D of this step's u
            D.eval_grid_(outs_np[i % 2], itp_h, flags=1)    # kernel + D2H of this step's result
        itp_h.synchronize()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        e2e_sync_ms = None
        t0 = time.perf_counter()
        for i in range(3):
            itp_h.set_u(u_np[i % 2])         # fully synchronous calls, like the reference's eval_grid!
            D.eval_grid_(out_np, itp_h)
        e2e_sync_ms = (time.perf_counter() - t0) * 1e3 / 3
        esz = np.dtype(wl.dtype).itemsize
        h2d, d2h = int(np.prod(wl.n_u)) * esz, wl.points_per_step() * esz
        itp_h.close()
    else:
        n = out.shape[0]
        n_h = min(n, 100_000_000)     # pinned host buffers of up to 3.6 GB; a larger per-step point count is scaled (and says so)
        xs_h = [x[:n_h].cpu().pin_memory() for x in xs]
        out_h = torch.empty((int(np.prod(out.shape[1:])) if out.ndim > 1 else 1, n_h), dtype=tdt).pin_memory()
        out_np = out_h.numpy().T if out.ndim > 1 else out_h.numpy().reshape(-1)
        out_np = out_np.reshape((n_h,) + tuple(out.shape[1:]), order="F") if out.ndim > 1 else out_np
        pts = [x.numpy() for x in xs_h]
        for _ in range(2):
            itp.eval_points(pts, out=out_np)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            itp.eval_points(pts, out=out_np)   # H2D of the points + kernel + D2H of the result
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e * (n / n_h)   # scaled to the full per-step point count
        esz = np.dtype(wl.dtype).itemsize
        h2d, d2h = n * esz * len(pts), int(np.prod(out.shape)) * esz

    # max over ranks
    vals = torch.tensor([ms_total, e2e_ms or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(vals[0]), float(vals[1])
    ms_per_step = ms_total / args.steps
    pts_step_all = wl.points_per_step() * world
    value = pts_step_all / (ms_per_step * 1e-3) / 1e9
    peaks, peak_src = measured_peaks()
    traffic = ncu_traffic(wl.key + ("_sorted" if args.sorted else ""))
    achieved = wl.alg_bytes() / (ms_per_step * 1e-3) / 1e9     # per GPU: one launch per step per rank, the bytes that rank touches
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "Gpoints/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype_name, "data": "synthetic",
        "config": {"workload": wl.name, "points_per_step_per_gpu": wl.points_per_step(),
                   "sharding": (("slab along the last grid axis" if wl.grid else "contiguous index range") + ", u/knots replicated, no collective"
                                + (f"; weak scaling: the global grid's last axis is {world} sweeps over the knot span ({world}x the points), rank r "
                                   f"owns sweep r — every rank evaluates exactly the N = 1 problem (same u extent, density and algorithmic bytes)"
                                   if wl.grid and world > 1 else "")),
                   "l2": f"working set exceeds L2 (126 MB): output streamed, u rotates over {wl.n_u_buffers} buffers" if wl.grid else "t_eval + out exceed L2",
                   "kernel": kernel, "launch": launch_mode, "first_call_ms_incl_cache_build": round(first_call_ms, 3), "cache_build": cache_build, "gradient": args.gradient, "all_gather_outputs": bool(args.gather and world > 1), "point_order": ("iid random, evaluated through the cell-sorted plan" + (", output left in plan order" if args.plan_order else "") if args.sorted else "iid random") if not wl.grid else "grid"},
        "e2e": {"value": round(pts_step_all / (e2e_ms * 1e-3) / 1e9, 4), "unit": "Gpoints/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "host (pinned) buffers through the public API, H2D + kernel + D2H every step; grid workloads queue the steps (DIND_FLAG_ASYNC) so D2H(i) overlaps H2D(i+1)"
                        + ("" if wl.grid or wl.points_per_step() <= 100_000_000 else "; measured on 1e8 points per step and scaled to the step's point count (PCIe-bound, linear in n)"),
                "synchronous_calls_value": (round(pts_step_all / (e2e_sync_ms * 1e-3) / 1e9, 4) if wl.grid else None)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": (traffic or {}).get("bytes"),
                     "traffic_source": (traffic or {}).get("source"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": wl.alg_bytes()},
        "clocks": clocks.summary(),
    }
    itp.close()
    return line


# ---- the other BASELINE configs at their named sizes, in the same JSON line ---------------------------------------
FMA_PEAK_TFMA = {"f64": 17.6, "f32": 32.3}   # register-resident FMA microbenchmark on B200 (scripts/microbench_fma.cu; profiles/r2_microbench.txt)
L2_GATHER_GREQ_PER_S = 285.0                 # 32-byte requests per second (x 1e9) of a random gather from an L2-resident table, one lane per record
                                             # (scripts/microbench_gather.cu; profiles/r2b_microbench_gather.txt: 17.8 Grecords/s x 16 loads)
C3_FP64_OPS_PER_POINT = 378                  # 285 DFMA + 51 DADD + 42 DMUL per point and evaluation (ncu instruction mix)
C3_GRADIENT_FP64_OPS_PER_POINT = 756         # fused value + 3 partials: contraction 384 + 144 + 48 FMA, order-0 and order-1 basis ~180


def _time_steps(torch, dist, stream, world, fn, steps, warmup):
    """W warm-ups, then K steps between CUDA events on the launching stream; barrier + synchronize on both sides."""
    for i in range(warmup):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(warmup + i)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _max_over_ranks(torch, dist, world, dev, vals):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def _entry(wl, ms, n_points_all, alg_bytes_rank, launches, kernel, extra=None):
    peaks, _ = measured_peaks()
    achieved = alg_bytes_rank / (ms * 1e-3) / 1e9
    e = {"value": round(n_points_all / (ms * 1e-3) / 1e9, 3), "unit": "Gpoints/s", "ms_per_step": round(ms, 4),
         "roofline": {"bound": "hbm", "frac": round(achieved / peaks["hbm_gbs"], 4), "achieved": round(achieved, 1),
                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "algorithmic_bytes": int(alg_bytes_rank)},
         "kernel": kernel, "gpu_launches_per_step": launches, "dtype": wl.dtype_name}
    if extra:
        e.update(extra)
    return e


def bench_unstructured_family(D, torch, dist, wl, rank, world, local_rank, steps, warmup, modes, gather=False):
    """One point-sharded unstructured workload (weak scaling: wl.n_points per rank), several evaluation modes over the
    same device-resident points: iid order, through the cell-sorted plan (original order / plan order), fused gradient.
    Returns {mode: entry}.  With `gather`, also the step followed by the NCCL all-gather of the outputs."""
    dev = f"cuda:{local_rank}"
    stream = torch.cuda.current_stream()
    tdt = torch.float32 if wl.dtype == np.float32 else torch.float64
    wl.build(rank, world)
    n_total = wl.n_points * world
    lo, hi = D.shard_range(n_total, rank, world)
    n = hi - lo
    gen = torch.Generator(device=dev)
    gen.manual_seed(2000 + rank)
    xs = [torch.rand(n, generator=gen, device=dev, dtype=tdt) * float(t[-1] - t[0]) + float(t[0]) for t in wl.ts]
    mk = {"linear": lambda t, p, x: D.LinearInterpolationDimension(t, t_eval=x),
          "bspline": lambda t, p, x: D.BSplineInterpolationDimension(t, p, t_eval=x, max_derivative_order_eval=1)}
    dims = tuple(mk[k](t, p, x) for k, t, p, x in zip(wl.kinds, wl.ts, wl.degrees, xs))
    shape = tuple(wl.n_u)
    n_in = len(wl.kinds)
    ud = D.f_empty(shape, wl.dtype, dev)
    gen_u = torch.Generator(device=dev)
    gen_u.manual_seed(77)                       # u replicated: the same table on every rank
    ud.copy_(torch.rand(shape[::-1], generator=gen_u, device=dev, dtype=tdt).permute(*reversed(range(len(shape)))))
    itp = D.NDInterpolation(ud, dims, device=local_rank, stream=stream.cuda_stream)
    out = D.f_empty((n,) + shape[n_in:], wl.dtype, dev)
    esz = np.dtype(wl.dtype).itemsize
    n_comp = int(np.prod(shape[n_in:])) if len(shape) > n_in else 1
    u_bytes = int(np.prod(shape)) * esz
    res = {}
    FL = {"iid": 1, "sorted": 1 | 16, "plan_order": 1 | 16 | 32}
    for mode in modes:
        extra = {"points_per_gpu": n, "point_order": {"iid": "iid random", "sorted": "iid random, evaluated through the cell-sorted plan, results in original order",
                                                      "plan_order": "iid random, evaluated through the cell-sorted plan, results left in plan order",
                                                      "gradient": "value + all first partials in one pass through the cell-sorted plan, results in original order",
                                                      "gradient_plan_order": "value + all first partials in one pass through the cell-sorted plan, results left in plan order"}[mode]}
        if mode.startswith("gradient"):
            gout = D.f_empty((n,) + shape[n_in:] + (n_in + 1,), wl.dtype, dev)
            fn = lambda i, f=(1 | 16 | (32 if mode.endswith("plan_order") else 0)): D.eval_unstructured_gradient(itp, gout, flags=f)
            alg = n * (n_in + n_comp * (n_in + 1)) * esz + u_bytes      # 120 B/pt for config 3
        else:
            fn = lambda i, f=FL[mode]: D.eval_unstructured_(out, itp, flags=f)
            alg = n * (n_in + n_comp) * esz + u_bytes
        if mode != "iid":
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            itp.build_caches(grid=False, flags=1 | 16)
            torch.cuda.synchronize()
            extra["plan_build_ms_first"] = round((time.perf_counter() - t0) * 1e3, 3)
        fn(0)                       # first call: one-time work (component-fastest copy of u, work items of the plan)
        l0 = itp.launch_count
        fn(1)
        per_step = itp.launch_count - l0
        ms = _time_steps(torch, dist, stream, world, fn, steps, warmup)
        kernel = itp.last_kernel
        (ms,) = _max_over_ranks(torch, dist, world, dev, [ms])
        if wl.key == "c3":
            ops = C3_GRADIENT_FP64_OPS_PER_POINT if mode.startswith("gradient") else C3_FP64_OPS_PER_POINT
            fma_bound = FMA_PEAK_TFMA["f64"] * 1e12 / ops / 1e9   # Gpoints/s per GPU
            extra["roofline_fma"] = {"bound": "fp64 pipe", "peak_Gpoints_per_s_per_gpu": round(fma_bound, 1),
                                     "frac": round(n / (ms * 1e-3) / 1e9 / fma_bound, 4),
                                     "note": f"{ops} FP64 operations per point at the measured 17.6 T/s (DESIGN.md §6)"}
        if wl.key == "c5" and mode == "iid":
            # iid order: no two points share a cell, so every point must fetch its own 2^5 corner nodes (512 B) from a
            # table that does not fit in L2 — the DRAM traffic that is necessary in this order (ncu: 523 B/point read)
            need = n * (2 ** n_in * n_comp * esz + (n_in + n_comp) * esz)
            gbs = need / (ms * 1e-3) / 1e9
            hbm = measured_peaks()[0]["hbm_gbs"]
            extra["roofline_gather"] = {"bound": "hbm, iid point order", "bytes_per_point": need // n, "achieved": round(gbs, 1),
                                        "peak": hbm, "unit": "GB/s", "frac": round(gbs / hbm, 4),
                                        "note": "cell-record copy of u (7.6 GB): one contiguous 512-byte block per point, requested as whole "
                                                "128-byte lines by four lanes (dind_fast_unstructured_rec.cu)"}
        if wl.key == "c3" and mode == "iid":
            # iid order: every lane fetches its own 4 x 4 x 4 stencil from the L2-resident component-fastest copy (67 MB) as
            # 64 requests of one 32-byte node; an SM completes about one such request per clock (the gather microbenchmark on
            # an L2-resident table, one lane per record, 32-byte loads: profiles/r2b_microbench_gather.txt)
            req = L2_GATHER_GREQ_PER_S
            bound = req / 64.0
            extra["roofline_gather"] = {"bound": "L1 miss requests of an L2-resident gather, iid point order", "requests_per_point": 64,
                                        "peak_Grequests_per_s": req, "peak_Gpoints_per_s_per_gpu": round(bound, 2),
                                        "frac": round(n / (ms * 1e-3) / 1e9 / bound, 4),
                                        "note": "scripts/microbench_gather.cu, table of 64 MB, 32-byte loads, one lane per record"}
        res[mode] = _entry(wl, ms, n_total, alg, per_step, kernel, extra)
        if mode.startswith("gradient"):
            del gout
    if gather and world > 1:
        # the optional NCCL all-gather of the outputs (dind_allgather_out), after the kernel, never inside it
        full = D.gather_unstructured(out, n_total, interp=itp)     # sets up the communicator
        del full
        torch.cuda.empty_cache()
        fullbuf = [None]

        def fn_g(i):
            D.eval_unstructured_(out, itp, flags=1)
            fullbuf[0] = None
            fullbuf[0] = D.gather_unstructured(out, n_total, interp=itp)
        ms_g = _time_steps(torch, dist, stream, world, fn_g, max(2, min(steps, 5)), 2)
        fn_e = lambda i: D.eval_unstructured_(out, itp, flags=1)
        ms_e = _time_steps(torch, dist, stream, world, fn_e, max(2, min(steps, 5)), 1)
        ms_g, ms_e = _max_over_ranks(torch, dist, world, dev, [ms_g, ms_e])
        recv = (n_total - n) * n_comp * esz
        res["iid"]["with_allgather"] = {
            "value": round(n_total / (ms_g * 1e-3) / 1e9, 3), "ms_per_step": round(ms_g, 3),
            "allgather_ms": round(ms_g - ms_e, 3), "bytes_received_per_gpu": int(recv),
            "busbw_GBps": round(recv / max(ms_g - ms_e, 1e-6) / 1e6, 1),
            "what": f"dind_allgather_out: {n_comp} ncclAllGather calls (one per component column) in one NCCL group, straight into the "
                    f"(n_total, {n_comp}) result; NVLink/NVSwitch bound"}
        fullbuf[0] = None
        torch.cuda.empty_cache()
        # the same work as ONE collective call: the shard is evaluated straight into its rows of the global array and the
        # rows of the other ranks arrive chunk by chunk while the next chunk is computed (dind_eval_unstructured_allgather)
        full2 = D.f_empty((n_total,) + shape[n_in:], wl.dtype, dev)
        fn_f = lambda i: D.eval_unstructured_allgather_(full2, itp, n_total)
        ms_f = _time_steps(torch, dist, stream, world, fn_f, max(2, min(steps, 5)), 2)
        (ms_f,) = _max_over_ranks(torch, dist, world, dev, [ms_f])
        res["iid"]["with_allgather_overlapped"] = {
            "value": round(n_total / (ms_f * 1e-3) / 1e9, 3), "ms_per_step": round(ms_f, 3),
            "what": "dind_eval_unstructured_allgather: no local output array; NCCL send/recv groups of chunk j on the communicator's "
                    "stream under the kernel of chunk j + 1; bound by max(evaluation, exchange) instead of their sum"}
        del full2
    itp.close()
    del xs, out, ud
    torch.cuda.empty_cache()
    return res


def bench_grid_config(D, torch, dist, wl, rank, world, local_rank, steps, warmup):
    dev = f"cuda:{local_rank}"
    stream = torch.cuda.current_stream()
    tdt = torch.float32 if wl.dtype == np.float32 else torch.float64
    wl.build(rank, world)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3000 + rank)
    rev = tuple(reversed(range(len(wl.n_u))))

    def to_dev(a):
        t = D.f_empty(a.shape, wl.dtype, dev)
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(dev))
        return t
    dims = wl.make_dims(D, [to_dev(x) for x in wl.t_evals])
    u_bufs = []
    for _ in range(wl.n_u_buffers):
        ub = D.f_empty(wl.n_u, wl.dtype, dev)
        ub.copy_(torch.rand(wl.n_u[::-1], generator=gen, device=dev, dtype=tdt).permute(*rev))
        u_bufs.append(ub)
    cache = None
    if wl.nurbs:
        wts = D.f_empty(wl.n_u, wl.dtype, dev)
        wts.copy_((torch.rand(wl.n_u[::-1], generator=gen, device=dev, dtype=tdt) + 0.5).permute(*rev))
        cache = D.NURBSWeights(wts)
    itp = D.NDInterpolation(u_bufs[0], dims, cache=cache, device=local_rank, stream=stream.cuda_stream)
    out = D.f_empty(wl.n_grid, wl.dtype, dev)

    def fn(i):
        itp.set_u(u_bufs[i % len(u_bufs)])
        D.eval_grid_(out, itp, flags=1)
    fn(0)
    l0 = itp.launch_count
    fn(1)
    per_step = itp.launch_count - l0
    ms = _time_steps(torch, dist, stream, world, fn, steps, warmup)
    (ms,) = _max_over_ranks(torch, dist, world, dev, [ms])
    e = _entry(wl, ms, wl.points_per_step() * world, wl.alg_bytes(), per_step, itp.last_kernel,
               {"grid": list(wl.n_grid), "u": list(wl.n_u)})
    itp.close()
    del out, u_bufs
    torch.cuda.empty_cache()
    return e


def run_all_configs(args, rank, world, local_rank):
    """Every BASELINE config other than the headline, at its named size (module docstring)."""
    import torch
    import torch.distributed as dist

    import dind_b200 as D
    K, W = max(3, min(args.steps, 10)), 3
    if world == 1:
        cfg = {}
        r = bench_unstructured_family(D, torch, dist, C1(), rank, world, local_rank, max(K, 20), 5, ["iid"])
        cfg["c1"] = r["iid"]
        r = bench_unstructured_family(D, torch, dist, C3(100_000_000), rank, world, local_rank, K, W, ["iid", "sorted", "plan_order", "gradient", "gradient_plan_order"])
        cfg.update({"c3": r["iid"], "c3_sorted": r["sorted"], "c3_plan_order": r["plan_order"], "c3_gradient": r["gradient"],
                    "c3_gradient_plan_order": r["gradient_plan_order"]})
        cfg["c4"] = bench_grid_config(D, torch, dist, C4(), rank, world, local_rank, K, W)
        cfg["c4b"] = bench_grid_config(D, torch, dist, C4b(), rank, world, local_rank, K, W)
        r = bench_unstructured_family(D, torch, dist, C5(500_000_000), rank, world, local_rank, max(3, K // 2), 2, ["iid"])
        cfg["c5"] = r["iid"]
        r = bench_unstructured_family(D, torch, dist, C5(100_000_000), rank, world, local_rank, K, W, ["sorted", "plan_order"])
        cfg.update({"c5_sorted": r["sorted"], "c5_plan_order": r["plan_order"]})
        for k, v in cfg.items():
            t = ncu_traffic(k)
            v["roofline"]["traffic"] = t["bytes"] if t else None
            if t:
                v["roofline"]["traffic_source"] = t["source"]
        return "configs", cfg
    ps = {}
    r = bench_unstructured_family(D, torch, dist, C5(500_000_000), rank, world, local_rank, max(3, K // 2), 2, ["iid"], gather=True)
    ps["c5"] = r["iid"]
    ps["c5"]["note"] = (f"{world} x 5e8 points sharded by contiguous index range (at 8 GPUs: BASELINE configs[4], 4e9 points), u (32^5, 4) replicated, "
                        "no collective in the evaluation; `with_allgather` adds dind_allgather_out of the (n_total, 4) Float32 output")
    r = bench_unstructured_family(D, torch, dist, C5(100_000_000), rank, world, local_rank, K, W, ["sorted", "plan_order"])
    ps["c5_sorted"], ps["c5_plan_order"] = r["sorted"], r["plan_order"]
    r = bench_unstructured_family(D, torch, dist, C3(100_000_000), rank, world, local_rank, K, W, ["iid", "sorted", "plan_order"], gather=True)
    ps["c3"], ps["c3_sorted"], ps["c3_plan_order"] = r["iid"], r["sorted"], r["plan_order"]
    return "point_sharded", ps


def bind_to_gpu_cpus(local_rank):
    """Pin this rank to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root), BEFORE any pinned host
    buffer is allocated: the e2e leg streams 600 MB per step through host memory, and with 8 ranks on a two-socket box
    remote-socket buffers halve the copy bandwidth.  Best effort: returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"gpu-local cpus ({len(allowed)} of {n_cpu})"
        return "unchanged (no gpu-local cpu in the allowed set)"
    except Exception as e:   # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- CPU arm (oracle port of the reference) -----------------------------------------------------------
def cpu_sample(wl, seconds_target=12.0, nthreads=0):
    """Times the CPU restatement of the reference on a bounded sample of the workload."""
    from oracle import oracle as O
    wl.build(0, 1)
    rng = np.random.default_rng(5)
    nthreads = nthreads or host_threads()
    if wl.grid:
        u = rng.random(wl.n_u, dtype=np.float32).astype(wl.dtype)
        u = np.asfortranarray(u)
        wts = np.asfortranarray(0.5 + rng.random(wl.n_u, dtype=np.float32).astype(wl.dtype)) if wl.nurbs else None
        k = 2
        sample = None
        while True:
            te = list(wl.t_evals[:-1]) + [wl.t_evals[-1][:k]]
            t0 = time.perf_counter()
            O.evaluate(list(wl.kinds), wl.ts, u, te, degrees=list(wl.degrees), weights=wts, grid=True, nthreads=nthreads)
            dt = time.perf_counter() - t0
            pts = int(np.prod([len(x) for x in te]))
            sample = (pts, O.last_eval_seconds(), f"{'x'.join(str(len(x)) for x in te)} slab of the {wl.n_grid} grid, full u; eval loop only (idx cache prebuilt, as the reference does at construction)")
            if dt > seconds_target / 3 or k >= len(wl.t_evals[-1]):
                break
            k = min(len(wl.t_evals[-1]), max(k * 2, int(k * seconds_target / 2 / max(dt, 1e-3))))
        return sample[0] / sample[1] / 1e9, nthreads, sample[2]
    shape = tuple(wl.n_u)
    u = np.asfortranarray(rng.random(shape).astype(wl.dtype))
    n = 200_000
    while True:
        xs = [rng.uniform(t[0], t[-1], n).astype(wl.dtype) for t in wl.ts]
        t0 = time.perf_counter()
        O.evaluate(list(wl.kinds), wl.ts, u, xs, degrees=list(wl.degrees), nthreads=nthreads)
        dt = time.perf_counter() - t0
        if dt > seconds_target / 3 or n >= wl.n_points:
            return n / O.last_eval_seconds() / 1e9, nthreads, f"{n} of {wl.n_points} points, full u; eval loop only (idx/basis caches prebuilt, as the reference does at construction)"
        n = min(wl.n_points, max(n * 2, int(n * seconds_target / 2 / max(dt, 1e-3))))


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    nthreads = host_threads()   # all host cores, whatever OMP_NUM_THREADS torchrun exported
    from oracle import oracle as O
    vals, secs = [], []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, _, sample = cpu_sample(wl, seconds_target=max(3.0, 60.0 / (args.warmup + args.steps)), nthreads=nthreads)
        if i >= args.warmup:
            vals.append(v)
            secs.append(O.last_eval_seconds())   # the step = one evaluation of the bounded sample
    value = float(np.mean(vals))
    return {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "Gpoints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * float(np.mean(secs)), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype_name, "data": "synthetic",
        "config": {"workload": wl.name, "note": "CPU restatement of the reference (oracle/, C++/OpenMP), not Julia: Julia is not in this image"},
        "cpu_baseline": {"value": round(value, 5), "unit": "Gpoints/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 5), "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def make_workload(args):
    if args.workload == "c2":
        return C2(args.scale)
    if args.workload == "c4":
        return C4(args.scale)
    if args.workload == "c4b":
        return C4b(args.scale)
    if args.workload == "c1":
        return C1()
    if args.workload == "c3":
        return C3(args.points or (100_000_000 if args.full else 20_000_000))
    if args.workload == "c5":
        return C5(args.points or (4_000_000_000 if args.full else 100_000_000))
    raise SystemExit(f"unknown workload {args.workload}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="dind")
    ap.add_argument("--full", action="store_true", help="full-size point counts for c3 / c5")
    ap.add_argument("--points", type=int, default=0, help="c3 / c5: points per GPU (overrides --full)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink c2 (debugging only; not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="grid headline: time the eager Python loop instead of CUDA-graph replays")
    ap.add_argument("--only-headline", action="store_true", help="skip the other BASELINE configs (`configs` / `point_sharded`)")
    ap.add_argument("--no-cache-build", action="store_true", help="skip the separate timing of cache construction")
    ap.add_argument("--gradient", choices=["fused", "separate"], default=None,
                    help="unstructured workloads: a step evaluates value + all first partials (one fused pass, or N_in+1 calls)")
    ap.add_argument("--gather", action="store_true", help="N > 1: all-gather the outputs (NCCL) inside the timed step")
    ap.add_argument("--plan-order", action="store_true", help="with --sorted: leave the output in plan order (no scatter)")
    ap.add_argument("--sorted", action="store_true", help="unstructured workloads: evaluate through the cell-sorted plan")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    wl = make_workload(args)

    if args.impl == "reference":
        line = run_reference(args, wl, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    affinity = None
    if world > 1:
        affinity = bind_to_gpu_cpus(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    line = run_gpu(args, wl, rank, world, local_rank)
    line["config"]["cpu_affinity"] = affinity
    if not args.only_headline and args.workload == "c2":
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        key, val = run_all_configs(args, rank, world, local_rank)
        line[key] = val
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_sample(make_workload(args))
            line["cpu_baseline"] = {"value": round(v, 5), "unit": "Gpoints/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Batch drivers: parameter sweeps as independent points on the engine.

  * :func:`sweep_velocity`  <- ``sweep_velocity``  (/root/reference/scripts/sweep_superluminal.py:34-121)
  * :func:`sweep_sigma_z0`  <- ``sweep_sigma_z0``  (/root/reference/src/irrotational_warp/sweep.py:30-58)
  * :func:`sweep_2d_z0`     <- ``sweep_2d_z0``     (/root/reference/src/irrotational_warp/sweep.py:61-103)

The reference evaluates the points in a serial Python loop.  Here a 3D velocity sweep is ONE call
to ``iwb_energy3d_batch`` (several evaluations in flight, one host synchronisation), and under
``torch.distributed`` the points are dealt round-robin to the ranks (point q -> rank q mod P) with a
single all-gather of the small result records at the end.  Output dictionaries keep the reference's
keys.
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from .backend import (BACKEND_NAME, EPS, _grid, compute_energy_axisymmetric, two_point_tail_extrapolation)
from .engine import get_engine
from .slice2d import compute_slice_z0, integrate_signed


def partition_points(npoints: int, rank: int, world_size: int) -> list[int]:
    """Indices of the points rank ``rank`` evaluates (round robin)."""
    return list(range(rank, npoints, world_size))


def _sweep_record(v, e_pos_r1, e_neg_r1, e_pos_r2, e_neg_r2, r1, r2, elapsed):
    e_pos_inf = two_point_tail_extrapolation(e_pos_r1, e_pos_r2, r1, r2)
    e_neg_inf = two_point_tail_extrapolation(e_neg_r1, e_neg_r2, r1, r2)
    e_abs_inf = abs(e_pos_inf) + abs(e_neg_inf)
    e_net_inf = e_pos_inf + e_neg_inf
    return {
        "v": v,
        "e_pos_inf": e_pos_inf,
        "e_neg_inf": e_neg_inf,
        "e_net_inf": e_net_inf,
        "e_abs_inf": e_abs_inf,
        "tail_net_frac": abs(e_net_inf) / e_abs_inf if e_abs_inf > 0 else 0.0,
        "ratio_r2": e_pos_r2 / abs(e_neg_r2) if e_neg_r2 != 0 else None,
        "timing_s": elapsed,
    }


def sweep_velocity(*, mode: str, rho: float, sigma: float, v_values: list, extent: float,
                   tail_r1_mult: float, tail_r2_mult: float, backend: str = BACKEND_NAME, dtype: str = "float64",
                   nx: int = None, ny: int = None, n: int = None, group=None, distributed: bool = False) -> dict:
    """Velocity sweep; same result dictionary as the reference (sweep_superluminal.py:97-121).

    By default every calling process evaluates every point, like the reference.  With ``distributed=True`` (or an
    explicit ``group``) the call is COLLECTIVE: every rank of the group must make it with the same arguments; the
    points are dealt round-robin to the ranks and the records are all-gathered, so every rank returns the full
    result.  A process that merely has ``torch.distributed`` initialised is never partitioned without asking.
    """
    if backend != BACKEND_NAME:
        raise ValueError(f"Unknown backend: {backend!r} (expected: b200)")
    v_values = [float(v) for v in v_values]
    if not v_values:
        raise ValueError("v_values is empty")      # the reference dies with UnboundLocalError here
    r1, r2 = tail_r1_mult * rho, tail_r2_mult * rho
    rank, world = 0, 1
    if group is not None or distributed:
        if not _dist_ready():
            raise RuntimeError("sweep_velocity(distributed=True) needs an initialised torch.distributed process group")
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = partition_points(len(v_values), rank, world)
    records = {}
    if mode == "axisym":
        grid = None
        for q in mine:
            t0 = time.perf_counter()
            run = compute_energy_axisymmetric(rho=rho, sigma=sigma, v=v_values[q], extent=extent, nx=nx, ny=ny,
                                              backend=backend, dtype=dtype, shells=[r1, r2])
            e_at_r = run.pop("_e_at_r")
            (p1, n1), (p2, n2) = e_at_r(r1), e_at_r(r2)
            records[q] = _sweep_record(v_values[q], p1, n1, p2, n2, r1, r2, time.perf_counter() - t0)
            grid = run["grid"]
        if grid is None:
            x, dx = _grid(-extent, extent, nx, dtype)
            y, dy = _grid(0.0, extent, ny, dtype)
            grid = {"nx": nx, "ny": ny, "extent": extent, "dx": dx, "dy": dy}
    else:
        x, dx = _grid(-extent, extent, n, dtype)
        grid = {"n": n, "extent": extent, "dx": dx, "dy": dx, "dz": dx}
        t0 = time.perf_counter()
        pts = [dict(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=rho, sigma=sigma, v=v_values[q], dtype=dtype,
                    shells=[r1, r2], eps=EPS) for q in mine]
        results = get_engine().energy3d_batch(pts) if pts else []
        elapsed = (time.perf_counter() - t0) / max(len(mine), 1)
        shell_sums = {q: [res["shell_pos"][0], res["shell_neg"][0], res["shell_pos"][1], res["shell_neg"][1], elapsed]
                      for q, res in zip(mine, results)}
        if world > 1:       # one all-gather of 5 doubles per point; every rank then forms every record
            shell_sums = _gather_values(shell_sums, len(v_values), 5, group)
        for q, (p1, n1, p2, n2, dt) in shell_sums.items():
            records[q] = _sweep_record(v_values[q], p1, n1, p2, n2, r1, r2, dt)
    if world > 1 and mode == "axisym":
        records = _gather_records(records, group)
    return {
        "sweep_results": [records[q] for q in range(len(v_values))],
        "params": {"mode": mode, "rho": rho, "sigma": sigma, "extent": extent, "r1": r1, "r2": r2,
                   "backend": backend, "dtype": dtype},
        "grid": grid,
    }


def _dist_ready() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def _gather_values(values: dict, npoints: int, width: int, group=None) -> dict:
    """``{point index: [width floats]}`` of this rank's round-robin share -> the same for all points, on every rank, with
    one ``all_gather_into_tensor`` (device tensors under NCCL, host tensors under gloo)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = max(-(-npoints // world), 1)
    local = torch.full((rows, width), float("nan"), dtype=torch.float64)
    for q, vals in values.items():
        local[q // world] = torch.tensor(vals, dtype=torch.float64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = local.to(dev)
    full = torch.empty((world, rows, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full.view(-1), local.view(-1), group=group)
    full = full.cpu()
    return {q: [float(t) for t in full[q % world, q // world]] for q in range(npoints)}


def _gather_records(records: dict, group=None) -> dict:
    import torch.distributed as dist
    world = dist.get_world_size(group)
    box = [None] * world
    dist.all_gather_object(box, records, group=group)
    merged = {}
    for part in box:
        merged.update(part)
    return merged


# ---- 2D-slice sweeps (package API) ---------------------------------------------------------------
@dataclass(frozen=True)
class SweepPoint:
    sigma: float
    e_pos: float
    e_neg: float
    e_net: float
    neg_fraction: float


@dataclass(frozen=True)
class SweepPoint2D:
    sigma: float
    v: float
    e_pos: float
    e_neg: float
    e_net: float
    neg_fraction: float


def _slice_point(rho, sigma, v, extent, n):
    res = compute_slice_z0(rho=rho, sigma=float(sigma), v=float(v), extent=extent, n=n)
    e_pos, e_neg, e_net = integrate_signed(res, dx=res.dx, dy=res.dy)
    denom = e_pos + e_neg
    return float(e_pos), float(e_neg), float(e_net), float((e_neg / denom) if denom > 0.0 else 0.0)


def sweep_sigma_z0(*, rho: float, v: float, sigma_values, extent: float, n: int) -> list:
    out = []
    for sigma in sigma_values:
        e_pos, e_neg, e_net, frac = _slice_point(rho, sigma, v, extent, n)
        out.append(SweepPoint(sigma=float(sigma), e_pos=e_pos, e_neg=e_neg, e_net=e_net, neg_fraction=frac))
    return out


def sweep_2d_z0(*, rho: float, sigma_values, v_values, extent: float, n: int) -> list:
    out = []
    for sigma in sigma_values:
        for v in v_values:
            e_pos, e_neg, e_net, frac = _slice_point(rho, sigma, v, extent, n)
            out.append(SweepPoint2D(sigma=float(sigma), v=float(v), e_pos=e_pos, e_neg=e_neg, e_net=e_net,
                                    neg_fraction=frac))
    return out

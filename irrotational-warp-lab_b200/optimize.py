"""Objectives and the exhaustive search on the engine: mirror of ``irrotational_warp.optimize``.

  * :func:`objective_neg_energy`    <- ``objective_neg_energy`` (/root/reference/src/irrotational_warp/optimize.py:32-64)
  * :func:`grid_search`             <- ``grid_search``          (:67-131)
  * :func:`objective_neg_energy_3d` -- BASELINE config 5's extension: |E-| of the 3D exact-potential volume
  * :func:`evaluate_objective_3d_batch` -- the ``n_initial_points`` / grid batches of config 5 as one engine call

  * :func:`optimize_nelder_mead`    <- ``optimize_nelder_mead`` (:134-209)  SciPy simplex, penalty bounds
  * :func:`optimize_hybrid`         <- ``optimize_hybrid``      (:212-272)  grid search, then simplex
  * :func:`optimize_bayesian`       <- ``optimize_bayesian``    (:275-356)  scikit-optimize ``gp_minimize``

The optimisers themselves (SciPy Nelder-Mead, scikit-optimize ``gp_minimize``) stay on the host with
their sequential ask/tell order; only the objective runs on the engine.  Every driver takes
``objective="slice2d"`` (the reference's semantics) or ``objective="volume3d"`` (BASELINE config 5).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

try:                                     # optional, exactly as in the reference (optimize.py:9-14)
    from skopt import gp_minimize
    from skopt.space import Real
    SKOPT_AVAILABLE = True
except ImportError:
    SKOPT_AVAILABLE = False

from .backend import EPS, _grid
from .engine import get_engine
from .slice2d import compute_slice_z0, integrate_signed


@dataclass(frozen=True)
class OptimizationResult:
    best_params: dict
    best_value: float
    initial_params: dict
    initial_value: float
    n_evaluations: int
    success: bool
    method: str
    message: str


def objective_neg_energy(params, param_names, rho: float, extent: float, n: int) -> float:
    """|E-| of the 2D slice (already a positive magnitude from integrate_signed)."""
    param_dict = dict(zip(param_names, params))
    res = compute_slice_z0(rho=rho, sigma=param_dict.get("sigma", 3.0), v=param_dict.get("v", 1.5), extent=extent, n=n)
    _, e_neg, _ = integrate_signed(res, dx=res.dx, dy=res.dy)
    return float(e_neg)


def grid_search(*, rho: float, sigma_range, v_range, sigma_steps: int, v_steps: int, extent: float, n: int
                ) -> OptimizationResult:
    sigma_vals = np.linspace(sigma_range[0], sigma_range[1], sigma_steps)
    v_vals = np.linspace(v_range[0], v_range[1], v_steps)
    best_value, best_sigma, best_v, n_evals = float("inf"), sigma_vals[0], v_vals[0], 0
    for sigma in sigma_vals:
        for v in v_vals:
            value = objective_neg_energy(np.array([sigma, v]), ["sigma", "v"], rho, extent, n)
            n_evals += 1
            if value < best_value:
                best_value, best_sigma, best_v = value, sigma, v
    return OptimizationResult(
        best_params={"sigma": best_sigma, "v": best_v},
        best_value=best_value,
        initial_params={"sigma": sigma_vals[0], "v": v_vals[0]},
        initial_value=objective_neg_energy(np.array([sigma_vals[0], v_vals[0]]), ["sigma", "v"], rho, extent, n),
        n_evaluations=n_evals, success=True, method="grid_search", message=f"Evaluated {n_evals} grid points",
    )


def _points_3d(param_list, rho, extent, n, dtype):
    x, dx = _grid(-extent, extent, n, dtype)
    return [dict(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=rho, sigma=float(s), v=float(v), dtype=dtype,
                 shells=[], eps=EPS) for s, v in param_list]


def evaluate_objective_3d_batch(param_list, rho: float, extent: float, n: int, dtype: str = "float64",
                                distributed: bool = False, group=None) -> list:
    """|E-| of the 3D volume for a list of (sigma, v) points: one ``iwb_energy3d_batch`` call (several evaluations in
    flight on the device, one host synchronisation).

    With ``distributed=True`` (or an explicit ``group``) the call is COLLECTIVE over the ``torch.distributed`` group: the
    points are dealt round-robin to the ranks (point q -> rank q mod P), every rank evaluates its
    share on its own GPU, and one all-gather of the values gives every rank the full list -- the ``n_initial_points`` /
    grid batches of BASELINE config 5 spread over the GPUs of a node."""
    param_list = list(param_list)
    if not (distributed or group is not None):
        res = get_engine().energy3d_batch(_points_3d(param_list, rho, extent, n, dtype))
        return [abs(r["e_neg"]) for r in res]
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("evaluate_objective_3d_batch(distributed=True) needs an initialised torch.distributed process group")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = list(range(rank, len(param_list), world))
    res = get_engine().energy3d_batch(_points_3d([param_list[q] for q in mine], rho, extent, n, dtype)) if mine else []
    width = -(-len(param_list) // world) if param_list else 0
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    local = torch.full((max(width, 1),), float("nan"), dtype=torch.float64)
    for k, r in enumerate(res):
        local[k] = abs(r["e_neg"])
    local = local.to(dev)
    full = torch.empty((world, max(width, 1)), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full.view(-1), local, group=group)
    full = full.cpu()
    return [float(full[q % world, q // world]) for q in range(len(param_list))]


def objective_neg_energy_3d(params, param_names, rho: float, extent: float, n: int, dtype: str = "float64") -> float:
    """Drop-in objective with the 3D evaluator (same signature as :func:`objective_neg_energy`)."""
    d = dict(zip(param_names, params))
    return evaluate_objective_3d_batch([(d.get("sigma", 3.0), d.get("v", 1.5))], rho, extent, n, dtype)[0]


def _objective(kind: str):
    if kind == "slice2d":
        return objective_neg_energy
    if kind == "volume3d":
        return objective_neg_energy_3d
    raise ValueError(f"Unknown objective: {kind!r} (expected: slice2d|volume3d)")


def optimize_nelder_mead(*, rho: float, initial_sigma: float, initial_v: float, extent: float, n: int,
                         bounds: dict | None = None, objective: str = "slice2d") -> OptimizationResult:
    """Local simplex search from (initial_sigma, initial_v); points outside ``bounds`` cost 1e10 per violated bound
    (Nelder-Mead has no native bounds).  maxiter 100, xatol 1e-4, fatol 1e-6 as in the reference."""
    from scipy.optimize import minimize

    names = ["sigma", "v"]
    fun = _objective(objective)
    s_lo, s_hi = (bounds or {}).get("sigma", (0.1, 100.0))
    v_lo, v_hi = (bounds or {}).get("v", (0.01, 10.0))

    def cost(params):
        if bounds is not None:
            sigma, v = params
            penalty = (1e10 if (sigma < s_lo or sigma > s_hi) else 0.0) + (1e10 if (v < v_lo or v > v_hi) else 0.0)
            if penalty > 0:
                return penalty
        return fun(params, names, rho, extent, n)

    n_evals = [0]

    def counted(params):
        n_evals[0] += 1
        return cost(params)

    x0 = np.array([initial_sigma, initial_v])
    initial_value = cost(x0)
    res = minimize(counted, x0, method="Nelder-Mead", options={"maxiter": 100, "xatol": 1e-4, "fatol": 1e-6})
    return OptimizationResult(best_params={"sigma": res.x[0], "v": res.x[1]}, best_value=res.fun,
                              initial_params={"sigma": initial_sigma, "v": initial_v}, initial_value=initial_value,
                              n_evaluations=n_evals[0], success=res.success, method="nelder_mead", message=res.message)


def optimize_hybrid(*, rho: float, sigma_range, v_range, sigma_steps: int, v_steps: int, extent: float, n: int,
                    refine: bool = True, objective: str = "slice2d") -> OptimizationResult:
    """Exhaustive grid, then Nelder-Mead from the best grid point inside the same bounds."""
    if objective == "slice2d":
        grid = grid_search(rho=rho, sigma_range=sigma_range, v_range=v_range, sigma_steps=sigma_steps, v_steps=v_steps,
                           extent=extent, n=n)
    else:       # the whole grid of the 3D objective is one batched engine call
        sig = np.linspace(sigma_range[0], sigma_range[1], sigma_steps)
        vel = np.linspace(v_range[0], v_range[1], v_steps)
        pts = [(s, v) for s in sig for v in vel]
        vals = evaluate_objective_3d_batch(pts, rho, extent, n)
        k = int(np.argmin(vals))          # first minimum, like the reference's strict "<" scan
        grid = OptimizationResult(best_params={"sigma": pts[k][0], "v": pts[k][1]}, best_value=vals[k],
                                  initial_params={"sigma": sig[0], "v": vel[0]}, initial_value=vals[0],
                                  n_evaluations=len(pts), success=True, method="grid_search",
                                  message=f"Evaluated {len(pts)} grid points")
    if not refine:
        return grid
    fine = optimize_nelder_mead(rho=rho, initial_sigma=grid.best_params["sigma"], initial_v=grid.best_params["v"],
                                extent=extent, n=n, bounds={"sigma": sigma_range, "v": v_range}, objective=objective)
    return OptimizationResult(best_params=fine.best_params, best_value=fine.best_value, initial_params=grid.initial_params,
                              initial_value=grid.initial_value, n_evaluations=grid.n_evaluations + fine.n_evaluations,
                              success=fine.success, method="hybrid_grid_nelder_mead",
                              message=f"Grid: {grid.n_evaluations} evals, Refinement: {fine.n_evaluations} evals")


def optimize_bayesian(*, rho: float, sigma_range, v_range, extent: float, n: int, n_calls: int = 50,
                      n_initial_points: int = 10, random_state=None, objective: str = "slice2d") -> OptimizationResult:
    """Gaussian-process search (EI acquisition, noise 1e-10) through scikit-optimize; ImportError without it."""
    if not SKOPT_AVAILABLE:
        raise ImportError("Bayesian optimization requires scikit-optimize. Install with: pip install scikit-optimize")
    fun = _objective(objective)
    n_evals = [0]

    def cost(params):
        n_evals[0] += 1
        return fun(np.array(params), ["sigma", "v"], rho, extent, n)

    res = gp_minimize(cost, [Real(sigma_range[0], sigma_range[1], name="sigma"), Real(v_range[0], v_range[1], name="v")],
                      n_calls=n_calls, n_initial_points=n_initial_points, random_state=random_state, acq_func="EI",
                      noise=1e-10, verbose=False)
    return OptimizationResult(best_params={"sigma": res.x[0], "v": res.x[1]}, best_value=res.fun,
                              initial_params={"sigma": res.x_iters[0][0], "v": res.x_iters[0][1]},
                              initial_value=float(res.func_vals[0]), n_evaluations=n_evals[0], success=True,
                              method="bayesian_gp",
                              message=f"Bayesian optimization with {n_calls} calls ({n_initial_points} initial random)")

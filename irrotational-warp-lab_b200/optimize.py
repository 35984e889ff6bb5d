"""Objectives and the exhaustive search on the engine: mirror of ``irrotational_warp.optimize``.

  * :func:`objective_neg_energy`    <- ``objective_neg_energy`` (/root/reference/src/irrotational_warp/optimize.py:32-64)
  * :func:`grid_search`             <- ``grid_search``          (:67-131)
  * :func:`objective_neg_energy_3d` -- BASELINE config 5's extension: |E-| of the 3D exact-potential volume
  * :func:`evaluate_objective_3d_batch` -- the ``n_initial_points`` / grid batches of config 5 as one engine call

The optimisers themselves (SciPy Nelder-Mead, scikit-optimize ``gp_minimize``) stay on the host with
their sequential ask/tell order; they take these objectives as plain callables.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

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


def evaluate_objective_3d_batch(param_list, rho: float, extent: float, n: int, dtype: str = "float64") -> list:
    """|E-| of the 3D volume for a list of (sigma, v) points: one ``iwb_energy3d_batch`` call."""
    res = get_engine().energy3d_batch(_points_3d(param_list, rho, extent, n, dtype))
    return [abs(r["e_neg"]) for r in res]


def objective_neg_energy_3d(params, param_names, rho: float, extent: float, n: int, dtype: str = "float64") -> float:
    """Drop-in objective with the 3D evaluator (same signature as :func:`objective_neg_energy`)."""
    d = dict(zip(param_names, params))
    return evaluate_objective_3d_batch([(d.get("sigma", 3.0), d.get("v", 1.5))], rho, extent, n, dtype)[0]

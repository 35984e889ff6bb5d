"""Far-field tail fit of a z=0 slice field: host-side post-process of the engine's rho_adm.

Mirror of ``irrotational_warp.tail`` (/root/reference/src/irrotational_warp/tail.py):

  * :class:`TailCorrectionResult`   <- ``TailCorrectionResult``        (:17-43)
  * :func:`radial_average_z0`       <- ``compute_radial_average_z0``   (:46-80)
  * :func:`fit_power_law_decay`     <- ``fit_power_law_decay``         (:83-131)
  * :func:`tail_integral_2d`        <- ``extrapolate_tail_integral_2d``(:134-161)
  * :func:`tail_uncertainty`        <- ``estimate_tail_uncertainty``   (:164-189)
  * :func:`compute_tail_correction` <- ``compute_tail_correction``     (:192-243)

SURVEY.md §2 row 8 / §8 row f3: this is an optional post-process of ``compute_slice_z0`` on a ~100^2 field
(50 radial bins, one ``polyfit``): microseconds of host work on data the slice evaluator has already returned,
so it stays on the host; the device part of the path is unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class TailCorrectionResult:
    r_bins: np.ndarray
    rho_avg: np.ndarray
    fit_r_min: float
    fit_r_max: float
    exponent: float
    amplitude: float
    tail_integral_pos: float
    tail_integral_neg: float
    tail_uncertainty: float
    fit_residual_rms: float


def radial_average_z0(field, x, y, n_bins: int = 50):
    """Mean of ``field`` over the annuli [r_k, r_k+1) of ``n_bins`` equal-width radial bins up to the
    farthest grid point (which, lying on the last closed edge, belongs to no bin -- as in the reference)."""
    X, Y = np.meshgrid(x, y)
    R = np.sqrt(X ** 2 + Y ** 2)
    edges = np.linspace(0, np.max(R), n_bins + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    avg = np.zeros(n_bins)
    for k in range(n_bins):
        sel = (R >= edges[k]) & (R < edges[k + 1])
        if np.sum(sel) > 0:
            avg[k] = np.mean(field[sel])
    return centres, avg


def fit_power_law_decay(r, rho, r_min: float, r_max: float):
    """Least-squares line through (log r, log|rho|) on r_min <= r <= r_max, |rho| > 1e-12:
    rho ~ A / r^n.  Returns (n, A, rms residual); (0, 0, inf) with fewer than three usable bins."""
    use = (r >= r_min) & (r <= r_max) & (np.abs(rho) > 1e-12)
    if np.sum(use) < 3:
        return 0.0, 0.0, np.inf
    rf, vf = r[use], rho[use]
    lr, lv = np.log(rf), np.log(np.abs(vf))
    slope, intercept = np.polyfit(lr, lv, deg=1)
    amplitude = np.exp(intercept)
    if np.mean(vf) < 0:
        amplitude = -amplitude
    resid = lv - (intercept + slope * lr)
    return float(-slope), float(amplitude), float(np.sqrt(np.mean(resid ** 2)))


def _tail_1d(exponent: float, r_start: float) -> float:
    # integral of r^(1-n) from R to infinity, n > 2
    return -r_start ** (2.0 - exponent) / (2.0 - exponent)


def tail_integral_2d(amplitude: float, exponent: float, r_start: float):
    """2 pi A * integral_R^inf r^(1-n) dr, split into (positive part, magnitude of negative part);
    (0, 0) when n <= 2 (the area integral does not converge)."""
    if exponent <= 2.0:
        return 0.0, 0.0
    tail = 2.0 * np.pi * amplitude * _tail_1d(exponent, r_start)
    return (float(tail), 0.0) if tail > 0 else (0.0, float(-tail))


def tail_uncertainty(amplitude: float, exponent: float, r_start: float, rms_residual: float) -> float:
    if exponent <= 2.0:
        return 0.0
    delta_a = np.abs(amplitude) * rms_residual
    return float(2.0 * np.pi * delta_a * np.abs(_tail_1d(exponent, r_start)))


def compute_tail_correction(field, x, y, fit_r_min=None, n_bins: int = 50) -> TailCorrectionResult:
    r_bins, rho_avg = radial_average_z0(field, x, y, n_bins=n_bins)
    r_max = r_bins[-1]
    if fit_r_min is None:
        fit_r_min = 0.7 * r_max
    exponent, amplitude, rms = fit_power_law_decay(r_bins, rho_avg, r_min=fit_r_min, r_max=r_max)
    pos, neg = tail_integral_2d(amplitude, exponent, r_start=r_max)
    return TailCorrectionResult(r_bins=r_bins, rho_avg=rho_avg, fit_r_min=fit_r_min, fit_r_max=r_max,
                                exponent=exponent, amplitude=amplitude, tail_integral_pos=pos, tail_integral_neg=neg,
                                tail_uncertainty=tail_uncertainty(amplitude, exponent, r_max, rms), fit_residual_rms=rms)


# ---- 3D generalisation on the engine ---------------------------------------------------------
# The reference fits the tail of the z = 0 slice only.  The same binning rule applied to the full 3D field of
# compute_energy_3d gives a shell-averaged profile <rho_adm>(r) whose power-law tail integrates to the energy beyond
# the box: E_tail = 4 pi A * integral_R^inf r^(2-n) dr.  The per-bin sums are produced on the device
# (csrc/radial_profile.cuh); nothing but n_bins numbers comes back.

def _tail_1d_3d(exponent: float, r_start: float) -> float:
    # integral of r^(2-n) from R to infinity, n > 3
    return -r_start ** (3.0 - exponent) / (3.0 - exponent)


def tail_integral_3d(amplitude: float, exponent: float, r_start: float):
    """4 pi A * integral_R^inf r^(2-n) dr as (positive part, magnitude of negative part); (0, 0) when n <= 3
    (the volume integral does not converge) -- the 3D twin of :func:`tail_integral_2d`."""
    if exponent <= 3.0:
        return 0.0, 0.0
    tail = 4.0 * np.pi * amplitude * _tail_1d_3d(exponent, r_start)
    return (float(tail), 0.0) if tail > 0 else (0.0, float(-tail))


def radial_profile_3d(*, rho: float, sigma: float, v: float, extent: float, n: int, n_bins: int = 50,
                      dtype: str = "float64", backend: str = "b200", mode: str = "exact") -> dict:
    """Shell average of the 3D rho_adm over ``n_bins`` equal-width radial bins from 0 to the farthest grid point
    (half-open bins, so the corner points belong to no bin -- the rule of :func:`radial_average_z0`)."""
    from .backend import EPS, _grid, resolve_backend
    eng, _ = resolve_backend(backend)
    x, dx = _grid(-extent, extent, n, dtype)
    xm = np.max(x * x)                                   # in the grid dtype, like X * X in the evaluator
    r_max = np.sqrt((xm + xm) + xm)                      # == np.max(np.sqrt(X*X + Y*Y + Z*Z)): rounding is monotone
    edges = np.linspace(0, r_max, n_bins + 1)            # a float32 grid gives float32 edges under NumPy >= 2
    res = eng.radial_profile3d(edges.astype(np.float64), x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=rho, sigma=sigma, v=v, dtype=dtype,
                               shells=[8.0 * rho, 12.0 * rho], eps=EPS, mode=mode)
    cnt, tot = res["bin_count"], res["bin_sum"]
    avg = np.zeros(n_bins)
    np.divide(tot, cnt, out=avg, where=cnt > 0)
    return {"r_bins": 0.5 * (edges[:-1] + edges[1:]), "rho_avg": avg, "edges": edges, "bin_sum": tot, "bin_count": cnt,
            "grid": {"n": n, "extent": extent, "dx": dx, "dy": dx, "dz": dx},
            "energies": {"e_pos": res["e_pos"], "e_neg": res["e_neg"], "e_net": float(res["e_pos"] + res["e_neg"])},
            "kernel_ms": res["kernel_ms"], "path": res.get("path")}


def compute_tail_correction_3d(*, rho: float, sigma: float, v: float, extent: float, n: int, fit_r_min=None,
                               fit_r_max=None, n_bins: int = 50, dtype: str = "float64",
                               backend: str = "b200") -> TailCorrectionResult:
    """:func:`compute_tail_correction` for the 3D field.  The fit window defaults to [0.7, 1] * extent: bins beyond
    the box half-width only see the corners of the cube, so they are left out of the fit and the tail starts at
    r = extent ... which the caller should read as "energy outside the inscribed sphere"."""
    prof = radial_profile_3d(rho=rho, sigma=sigma, v=v, extent=extent, n=n, n_bins=n_bins, dtype=dtype, backend=backend)
    r_bins, rho_avg = prof["r_bins"], prof["rho_avg"]
    r_hi = float(extent) if fit_r_max is None else float(fit_r_max)
    r_lo = 0.7 * r_hi if fit_r_min is None else float(fit_r_min)
    exponent, amplitude, rms = fit_power_law_decay(r_bins, rho_avg, r_min=r_lo, r_max=r_hi)
    pos, neg = tail_integral_3d(amplitude, exponent, r_start=r_hi)
    unc = 0.0 if exponent <= 3.0 else float(4.0 * np.pi * np.abs(amplitude) * rms * np.abs(_tail_1d_3d(exponent, r_hi)))
    return TailCorrectionResult(r_bins=r_bins, rho_avg=rho_avg, fit_r_min=r_lo, fit_r_max=r_hi, exponent=exponent,
                                amplitude=amplitude, tail_integral_pos=pos, tail_integral_neg=neg, tail_uncertainty=unc,
                                fit_residual_rms=rms)

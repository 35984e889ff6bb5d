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

"""2D z=0 slice diagnostic on the engine: mirror of ``irrotational_warp.adm``.

  * :class:`SliceResult`      <- ``adm.SliceResult``      (/root/reference/src/irrotational_warp/adm.py:14-24)
  * :func:`compute_slice_z0`  <- ``adm.compute_slice_z0`` (:27-88) + ``fd.central_diff_2d`` (fd.py:6-25)
  * :func:`integrate_signed`  <- ``adm.integrate_signed`` (:91-97)

The tanh-wall dipole potential, the edge_order=1 differences, K_xx/K_yy/K_xy and rho_ADM run in
``iwb_slice_z0`` (csrc/planar.cu).  ``tail_correction=True`` adds the reference's far-field fit as a host
post-process of the returned rho_adm (``tail.py``; 50 radial bins of a ~100^2 field).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .engine import get_engine


@dataclass(frozen=True)
class SliceResult:
    x: np.ndarray        # 1D
    y: np.ndarray        # 1D
    rho_adm: np.ndarray  # 2D [y,x]
    phi: np.ndarray      # 2D [y,x]
    beta_x: np.ndarray   # 2D [y,x]
    beta_y: np.ndarray   # 2D [y,x]
    dx: float
    dy: float
    tail_correction: Optional[object] = None
    # engine-side extras (already reduced on the device; integrate_signed uses them when present)
    signed_integrals: Optional[tuple] = None


def compute_slice_z0(*, rho: float, sigma: float, v: float, extent: float, n: int,
                     tail_correction: bool = False) -> SliceResult:
    """z=0 slice of phi, beta and rho_ADM = (K^2 - K_ij K^ij)/16 pi with the 2x2 K of the slice."""
    x = np.linspace(-extent, extent, n)
    y = np.linspace(-extent, extent, n)
    dx = float(x[1] - x[0])
    dy = float(y[1] - y[0])
    f = get_engine().slice_z0(x=x, y=y, dx=dx, dy=dy, rho=rho, sigma=sigma, v=v)
    tail = None
    if tail_correction:                                   # adm.py:73-75
        from .tail import compute_tail_correction
        tail = compute_tail_correction(f["rho_adm"], x, y, fit_r_min=0.7 * extent, n_bins=50)
    return SliceResult(x=x, y=y, rho_adm=f["rho_adm"], phi=f["phi"], beta_x=f["beta_x"], beta_y=f["beta_y"],
                       dx=dx, dy=dy, tail_correction=tail, signed_integrals=(f["pos"], f["neg"], f["net"]))


def integrate_signed(field, dx: float, dy: float):
    """(pos, neg, net) with ``neg`` a positive magnitude (adm.py:91-97).

    Accepts either a :class:`SliceResult` (uses the sums the device already reduced) or a plain
    array (reduced on the host exactly like the reference, for callers that post-process fields)."""
    if isinstance(field, SliceResult) and field.signed_integrals is not None:
        return field.signed_integrals
    arr = field.rho_adm if isinstance(field, SliceResult) else np.asarray(field)
    dA = dx * dy
    pos = float(np.sum(arr[arr > 0.0]) * dA)
    neg = float(np.sum(-arr[arr < 0.0]) * dA)
    return pos, neg, pos - neg

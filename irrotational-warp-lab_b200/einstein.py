"""Einstein-tensor eigenvalue diagnostic of the z = 0 slice on the engine.

Mirror of ``irrotational_warp.einstein`` (/root/reference/src/irrotational_warp/einstein.py):

  * :class:`EinsteinResult`               <- ``EinsteinResult``               (:23-38)
  * :func:`compute_einstein_eigenvalues`  <- ``compute_einstein_eigenvalues`` (:256-309), which chains
    ``compute_metric_z0`` (:41-71), ``compute_metric_inverse`` (:74-93), ``compute_christoffel`` (:96-164),
    ``compute_ricci_tensor`` (:167-226), ``compute_einstein_tensor`` (:229-253) and one ``np.linalg.eigvals``
    per grid point.

Same arguments and result object; the array work is two CUDA kernels (csrc/einstein.cu).  The eigenvalues in
``eig_all[:, i, j]`` are not in LAPACK's order (the reference's consumers use only ``eig_max``, ``ricci_scalar`` and
the Type-I fraction: viz.py:37-41, 98-103).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from .engine import get_engine

IMAG_TOL = 1e-8          # einstein.py:301


@dataclass(frozen=True)
class EinsteinResult:
    eig_max: np.ndarray        # [ny, nx] max real part of the eigenvalues of G^mu_nu
    eig_all: np.ndarray        # [4, ny, nx] complex
    ricci_scalar: np.ndarray   # [ny, nx]
    is_type_i: np.ndarray      # [ny, nx] bool: all eigenvalues real within IMAG_TOL


def compute_einstein_eigenvalues(beta_x: np.ndarray, beta_y: np.ndarray, dx: float, dy: float, *,
                                 device: int = 0) -> EinsteinResult:
    bx = np.ascontiguousarray(beta_x, dtype=np.float64)
    by = np.ascontiguousarray(beta_y, dtype=np.float64)
    if bx.ndim != 2 or bx.shape != by.shape:
        raise ValueError("beta_x and beta_y must be 2D arrays of the same shape [ny, nx]")
    ny, nx = bx.shape
    eng = get_engine(device)
    eig_re, eig_im = np.empty((4, ny, nx)), np.empty((4, ny, nx))
    eig_max, ricci = np.empty((ny, nx)), np.empty((ny, nx))
    type_i = np.empty((ny, nx), dtype=np.uint8)
    p = _capi.ParamsEinstein()
    p.nx, p.ny = nx, ny
    dbl = C.POINTER(C.c_double)
    p.beta_x, p.beta_y = bx.ctypes.data_as(dbl), by.ctypes.data_as(dbl)
    p.dx, p.dy, p.imag_tol = float(dx), float(dy), IMAG_TOL
    p.eig_re, p.eig_im = eig_re.ctypes.data_as(dbl), eig_im.ctypes.data_as(dbl)
    p.eig_max, p.ricci_scalar = eig_max.ctypes.data_as(dbl), ricci.ctypes.data_as(dbl)
    p.is_type_i = type_i.ctypes.data_as(C.POINTER(C.c_uint8))
    out = _capi.EinsteinResultC()
    _capi.check(eng._lib.iwb_einstein_z0(eng._h, C.byref(p), C.byref(out)))
    return EinsteinResult(eig_max=eig_max, eig_all=eig_re + 1j * eig_im, ricci_scalar=ricci,
                          is_type_i=type_i.astype(bool))

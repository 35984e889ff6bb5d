"""ctypes front-end of the C oracle (``oracle/oracle_c.c``).  TEST INFRASTRUCTURE ONLY.

Same contract as :mod:`oracle.oracle_np` (the NumPy restatement) but pointwise,
OpenMP-parallel and slab-able, so parity tests can afford n = 256..1024.
Coordinates, spacings and sinh/cosh(rho*sigma) are computed here with NumPy
exactly as the reference's host code does (reproduce_rodal_exact.py:171-177,
potential.py:50-52) and handed to C.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from .oracle_np import grid_axis

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_c.so")
_MAX_SHELLS = 8


class _Result(C.Structure):
    _fields_ = [("e_pos", C.c_double), ("e_neg", C.c_double),
                ("cnt_pos", C.c_int64), ("cnt_neg", C.c_int64), ("cnt_zero", C.c_int64),
                ("shell_pos", C.c_double * _MAX_SHELLS), ("shell_neg", C.c_double * _MAX_SHELLS),
                ("shell_cnt", C.c_int64 * _MAX_SHELLS)]


class _Params(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("v", C.c_double),
                ("sinh_rs", C.c_double), ("cosh_rs", C.c_double), ("dtype_mode", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_c.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.oracle_energy3d_slab.argtypes = [C.POINTER(_Params), C.c_int, C.c_int, C.c_int, dp, dp, dp,
                                              C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                              C.c_int, dp, dp, C.POINTER(_Result)]
        _lib.oracle_energy_axisym.argtypes = [C.POINTER(_Params), C.c_int, C.c_int, dp, dp,
                                              C.c_double, C.c_double, C.c_int, dp, dp, C.POINTER(_Result)]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _params(rho, sigma, v, dtype):
    rs = rho * sigma
    return _Params(rho, sigma, v, float(np.sinh(rs)), float(np.cosh(rs)), 1 if dtype == "float32" else 0)


def _unpack(res, shells):
    return {
        "e_pos": res.e_pos, "e_neg": res.e_neg,
        "cnt_pos": res.cnt_pos, "cnt_neg": res.cnt_neg, "cnt_zero": res.cnt_zero,
        "shells": [{"r": float(r), "e_pos": res.shell_pos[s], "e_neg": res.shell_neg[s],
                    "cnt": res.shell_cnt[s]} for s, r in enumerate(shells)],
    }


def energy_3d(*, rho, sigma, v, extent, n, dtype="float64", shells=(), i_begin=0, i_end=None,
              return_density=False):
    i_end = n if i_end is None else i_end
    x, dx = grid_axis(-extent, extent, n, dtype)
    xd = np.ascontiguousarray(x, dtype=np.float64)
    sh = np.ascontiguousarray(np.asarray(shells, dtype=np.float64))
    rho_out = np.empty((i_end - i_begin, n, n)) if return_density else None
    res = _Result()
    p = _params(rho, sigma, v, dtype)
    t0 = time.perf_counter()
    rc = lib().oracle_energy3d_slab(C.byref(p), n, n, n, _dp(xd), _dp(xd), _dp(xd), dx, dx, dx,
                                    i_begin, i_end, len(sh), _dp(sh), _dp(rho_out), C.byref(res))
    t1 = time.perf_counter()
    if rc == 1:
        raise ValueError("Shape of array too small to calculate a numerical gradient, "
                         "at least (edge_order + 1) elements are required.")
    if rc:
        raise RuntimeError(f"oracle_energy3d_slab failed rc={rc}")
    out = _unpack(res, sh)
    out.update(dx=dx, dy=dx, dz=dx, dV=dx * dx * dx, timing_s={"total": t1 - t0})
    if return_density:
        out["rho_adm"] = rho_out
    return out


def energy_3d_chunked(*, n, chunk=64, **kw):
    tot = None
    for a in range(0, n, chunk):
        part = energy_3d(n=n, i_begin=a, i_end=min(n, a + chunk), **kw)
        if tot is None:
            tot = part
            continue
        for key in ("e_pos", "e_neg", "cnt_pos", "cnt_neg", "cnt_zero"):
            tot[key] += part[key]
        for s_tot, s_part in zip(tot["shells"], part["shells"]):
            for key in ("e_pos", "e_neg", "cnt"):
                s_tot[key] += s_part[key]
        tot["timing_s"]["total"] += part["timing_s"]["total"]
    return tot


def energy_axisym(*, rho, sigma, v, extent, nx, ny, dtype="float64", shells=(), return_density=False):
    x, dx = grid_axis(-extent, extent, nx, dtype)
    y, dy = grid_axis(0.0, extent, ny, dtype)
    xd = np.ascontiguousarray(x, dtype=np.float64)
    yd = np.ascontiguousarray(y, dtype=np.float64)
    sh = np.ascontiguousarray(np.asarray(shells, dtype=np.float64))
    rho_out = np.empty((ny, nx)) if return_density else None
    res = _Result()
    p = _params(rho, sigma, v, dtype)
    t0 = time.perf_counter()
    rc = lib().oracle_energy_axisym(C.byref(p), nx, ny, _dp(xd), _dp(yd), dx, dy,
                                    len(sh), _dp(sh), _dp(rho_out), C.byref(res))
    t1 = time.perf_counter()
    if rc == 1:
        raise ValueError("Shape of array too small to calculate a numerical gradient, "
                         "at least (edge_order + 1) elements are required.")
    if rc:
        raise RuntimeError(f"oracle_energy_axisym failed rc={rc}")
    out = _unpack(res, sh)
    out.update(dx=dx, dy=dy, timing_s={"total": t1 - t0})
    if return_density:
        out["rho_adm"] = rho_out
    return out

"""CPU oracle (NumPy) for the energy-condition hot path.  TEST INFRASTRUCTURE ONLY.

This module is a from-scratch restatement, array-at-a-time, of the reference's
algorithm for the path named in BASELINE.json.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it; the product (``irrotational-warp-lab_b200``) never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` in the build container and freezes energies,
shell sums, sign counts and pointwise e-density digests into
``tests/golden/*.json``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors (bit-for-bit for the unchunked evaluation) and
``make_golden.py --check-oracle`` compares it pointwise with the live reference.

Reference lines followed (``/root/reference/...``):
  * ``src/irrotational_warp/potential.py:48-77``  g(r) (log-cosh via logaddexp)
  * ``src/irrotational_warp/potential.py:95-99``  phi = v r g cos(theta)
  * ``src/irrotational_warp/potential.py:11-12,31-34`` tanh-wall dipole (2D slice)
  * ``scripts/reproduce_rodal_exact.py:168-224``  3D evaluator
  * ``scripts/reproduce_rodal_exact.py:93-143``   axisymmetric evaluator
  * ``src/irrotational_warp/adm.py:47-71,91-97`` + ``fd.py:11-12``  2D z=0 slice
  * NumPy ``gradient`` (``numpy/lib/_function_base_impl.py``: interior
    ``(f[2:]-f[:-2])/(2.*h)``, edge_order 1 and 2 one-sided edges) -- restated by
    :func:`diff_axis` instead of being called, so that slabs of a larger volume
    can be differentiated without fake edges.

Every arithmetic step keeps the reference's operand order and dtype flow (the
reference's ``--dtype float32`` path is mixed precision under NumPy>=2: see
SURVEY.md fact 6); nothing here is "optimised".
"""
from __future__ import annotations

import math
import time

import numpy as np

_TOL_DEN = 1e-12  # potential.py:73
_EPS = 1e-12      # potential.py:88


# --------------------------------------------------------------------------- #
# finite differences (restating numpy.gradient for uniform spacing)
# --------------------------------------------------------------------------- #
def diff_axis(f, h, axis, *, edge_order=2, lo_edge=True, hi_edge=True):
    """d f / d(axis) with NumPy-gradient semantics for scalar spacing ``h``.

    ``lo_edge`` / ``hi_edge`` say whether the first / last plane of ``f`` along
    ``axis`` is a true face of the global grid (one-sided formula applies) or an
    artificial slab cut (the plane is filled with NaN and must not be used).
    """
    n = f.shape[axis]
    if n < edge_order + 1:
        raise ValueError(
            "Shape of array too small to calculate a numerical gradient, "
            "at least (edge_order + 1) elements are required."
        )
    out = np.empty_like(f)

    def sl(s):
        idx = [slice(None)] * f.ndim
        idx[axis] = s
        return tuple(idx)

    out[sl(slice(1, -1))] = (f[sl(slice(2, None))] - f[sl(slice(None, -2))]) / (2.0 * h)
    if edge_order == 1:
        lo = (f[sl(1)] - f[sl(0)]) / h
        hi = (f[sl(-1)] - f[sl(-2)]) / h
    else:
        a, b, c = -1.5 / h, 2.0 / h, -0.5 / h
        lo = a * f[sl(0)] + b * f[sl(1)] + c * f[sl(2)]
        a, b, c = 0.5 / h, -2.0 / h, 1.5 / h
        hi = a * f[sl(-3)] + b * f[sl(-2)] + c * f[sl(-1)]
    out[sl(0)] = lo if lo_edge else np.nan
    out[sl(-1)] = hi if hi_edge else np.nan
    return out


# --------------------------------------------------------------------------- #
# potentials
# --------------------------------------------------------------------------- #
def g_exact(r, rho, sigma):
    """Rodal's exact g(r); potential.py:48-77."""
    rs = rho * sigma
    s_rs = np.sinh(rs)      # np.float64 scalar: *strong* under NEP 50
    c_rs = np.cosh(rs)
    t1 = 2 * r * sigma * s_rs
    am = sigma * (r - rho)
    ap = sigma * (r + rho)
    lcm = np.logaddexp(am, -am)
    lcp = np.logaddexp(ap, -ap)
    t2 = c_rs * (lcm - lcp)
    num = t1 + t2
    den = 2 * r * sigma * s_rs
    ok = np.abs(den) > _TOL_DEN
    ratio = num / np.where(ok, den, 1.0)
    return np.where(ok, ratio, np.zeros_like(r))


def phi_exact(x, y, z, rho, sigma, v):
    """phi = v r g(r) x/(r+eps); potential.py:95-99."""
    r = np.sqrt(x * x + y * y + z * z)
    g = g_exact(r, rho, sigma)
    c = x / (r + _EPS)
    return v * r * g * c


def phi_dipole(x, y, z, rho, sigma, v):
    """tanh-wall dipole used by the 2D slice; potential.py:11-12,31-34."""
    r = np.sqrt(x * x + y * y + z * z)
    xi = r / rho
    f = 0.5 * (1.0 + np.tanh(sigma * (1.0 - xi)))
    c = x / (r + _EPS)
    return v * rho * f * c


# --------------------------------------------------------------------------- #
# 3D evaluator (optionally one x-slab of the global grid)
# --------------------------------------------------------------------------- #
def _np_dtype(dtype):
    return np.float32 if dtype == "float32" else np.float64


def grid_axis(lo, hi, n, dtype="float64"):
    """Coordinates and spacing exactly as the reference builds them
    (reproduce_rodal_exact.py:171-177): linspace then ``float(x[1]-x[0])``."""
    x = np.linspace(lo, hi, n, dtype=_np_dtype(dtype))
    return x, float(x[1] - x[0])


def energy_3d(*, rho, sigma, v, extent, n, dtype="float64", shells=(),
              i_begin=0, i_end=None, return_density=False):
    """3D energies on planes ``[i_begin, i_end)`` of the global n^3 grid.

    With the default full range this is reproduce_rodal_exact.py:168-224 step by
    step (and sums with one ``np.sum`` over the whole array like the reference).
    A sub-range evaluates phi on the slab plus a 2-plane overlap using the
    *global* coordinates, so partial results from disjoint ranges add up to the
    full answer; that is how n=1024^3 gets a CPU answer.
    """
    i_end = n if i_end is None else i_end
    x, dx = grid_axis(-extent, extent, n, dtype)
    y, dy = grid_axis(-extent, extent, n, dtype)
    z, dz = grid_axis(-extent, extent, n, dtype)

    lo = max(0, i_begin - 2)
    hi = min(n, i_end + 2)
    if lo == 0:
        hi = max(hi, min(n, 4))
    if hi == n:
        lo = min(lo, max(0, n - 4))
    X, Y, Z = np.meshgrid(x[lo:hi], y, z, indexing="ij")

    t0 = time.perf_counter()
    phi = phi_exact(X, Y, Z, rho, sigma, v)
    t_phi = time.perf_counter()
    edges = dict(lo_edge=(lo == 0), hi_edge=(hi == n))
    px = diff_axis(phi, dx, 0, **edges)
    py = diff_axis(phi, dy, 1)
    pz = diff_axis(phi, dz, 2)
    pxx = diff_axis(px, dx, 0, **edges)
    pxy = diff_axis(px, dy, 1)
    pxz = diff_axis(px, dz, 2)
    pyy = diff_axis(py, dy, 1)
    pyz = diff_axis(py, dz, 2)
    pzz = diff_axis(pz, dz, 2)
    del phi, px, py, pz

    kxx, kyy, kzz, kxy, kxz, kyz = -pxx, -pyy, -pzz, -pxy, -pxz, -pyz
    tr = kxx + kyy + kzz
    kk = kxx * kxx + kyy * kyy + kzz * kzz + 2.0 * (kxy * kxy + kxz * kxz + kyz * kyz)
    rho_adm = (tr * tr - kk) / (16.0 * np.pi)
    dV = dx * dy * dz
    e = rho_adm * dV

    own = slice(i_begin - lo, i_end - lo)
    e = e[own]
    r_grid = np.sqrt(X * X + Y * Y + Z * Z)[own]
    pos_m = e > 0.0
    neg_m = e < 0.0
    out = {
        "dx": dx, "dy": dy, "dz": dz, "dV": dV,
        "e_pos": float(np.sum(e * pos_m)),
        "e_neg": float(np.sum(e * neg_m)),
        "cnt_pos": int(np.count_nonzero(pos_m)),
        "cnt_neg": int(np.count_nonzero(neg_m)),
        "cnt_zero": int(np.count_nonzero(e == 0.0)),
        "shells": [],
    }
    for r_lim in shells:
        m = r_grid <= r_lim
        out["shells"].append({
            "r": float(r_lim),
            "e_pos": float(np.sum(e * pos_m * m)),
            "e_neg": float(np.sum(e * neg_m * m)),
            "cnt": int(np.count_nonzero(m)),
        })
    t1 = time.perf_counter()
    out["timing_s"] = {"phi": t_phi - t0, "total": t1 - t0}
    if return_density:
        out["e_density"] = e
    return out


def energy_3d_chunked(*, rho, sigma, v, extent, n, dtype="float64", shells=(), chunk=32):
    """Sum of :func:`energy_3d` over x-slabs of ``chunk`` planes (bounded memory)."""
    tot = None
    for a in range(0, n, chunk):
        part = energy_3d(rho=rho, sigma=sigma, v=v, extent=extent, n=n, dtype=dtype,
                         shells=shells, i_begin=a, i_end=min(n, a + chunk))
        if tot is None:
            tot = part
            continue
        for key in ("e_pos", "e_neg", "cnt_pos", "cnt_neg", "cnt_zero"):
            tot[key] += part[key]
        for s_tot, s_part in zip(tot["shells"], part["shells"]):
            for key in ("e_pos", "e_neg", "cnt"):
                s_tot[key] += s_part[key]
        for key in ("phi", "total"):
            tot["timing_s"][key] += part["timing_s"][key]
    return tot


# --------------------------------------------------------------------------- #
# axisymmetric evaluator
# --------------------------------------------------------------------------- #
def energy_axisym(*, rho, sigma, v, extent, nx, ny, dtype="float64", shells=(),
                  return_density=False):
    """reproduce_rodal_exact.py:93-143: meridional half-plane, arrays [ny, nx]."""
    x, dx = grid_axis(-extent, extent, nx, dtype)
    y, dy = grid_axis(0.0, extent, ny, dtype)
    X, Y = np.meshgrid(x, y, indexing="xy")
    Z = np.zeros_like(X)

    t0 = time.perf_counter()
    phi = phi_exact(X, Y, Z, rho, sigma, v)
    t_phi = time.perf_counter()
    py = diff_axis(phi, dy, 0)
    px = diff_axis(phi, dx, 1)
    pxx = diff_axis(px, dx, 1)
    pyy = diff_axis(py, dy, 0)
    pxy = diff_axis(py, dx, 1)          # d/dx of d(phi)/dy  (line 113)

    kxx, kyy, kxy = -pxx, -pyy, -pxy
    on_axis = Y == 0.0
    kzz = np.where(on_axis, kyy, (-py) / np.where(on_axis, 1.0, Y))
    kzz[0, :] = kyy[0, :]
    tr = kxx + kyy + kzz
    kk = kxx * kxx + kyy * kyy + kzz * kzz + 2.0 * kxy * kxy
    rho_adm = (tr * tr - kk) / (16.0 * np.pi)
    dV = 2.0 * np.pi * Y * dx * dy
    e = rho_adm * dV
    r_grid = np.sqrt(X * X + Y * Y)
    pos_m = e > 0.0
    neg_m = e < 0.0
    out = {
        "dx": dx, "dy": dy,
        "e_pos": float(np.sum(e * pos_m)),
        "e_neg": float(np.sum(e * neg_m)),
        "cnt_pos": int(np.count_nonzero(pos_m)),
        "cnt_neg": int(np.count_nonzero(neg_m)),
        "cnt_zero": int(np.count_nonzero(e == 0.0)),
        "shells": [],
    }
    for r_lim in shells:
        m = r_grid <= r_lim
        out["shells"].append({
            "r": float(r_lim),
            "e_pos": float(np.sum(e * pos_m * m)),
            "e_neg": float(np.sum(e * neg_m * m)),
            "cnt": int(np.count_nonzero(m)),
        })
    t1 = time.perf_counter()
    out["timing_s"] = {"phi": t_phi - t0, "total": t1 - t0}
    if return_density:
        out["e_density"] = e
    return out


# --------------------------------------------------------------------------- #
# 2D z=0 slice (dipole potential, edge_order=1)
# --------------------------------------------------------------------------- #
def slice_z0(*, rho, sigma, v, extent, n):
    """adm.py:47-71: returns dict of x, y, phi, beta_x, beta_y, rho_adm, dx, dy."""
    x = np.linspace(-extent, extent, n)
    y = np.linspace(-extent, extent, n)
    dx = float(x[1] - x[0])
    dy = float(y[1] - y[0])
    X, Y = np.meshgrid(x, y)
    phi = phi_dipole(X, Y, np.zeros_like(X), rho, sigma, v)
    bx = diff_axis(phi, dx, 1, edge_order=1)
    by = diff_axis(phi, dy, 0, edge_order=1)
    bx_x = diff_axis(bx, dx, 1, edge_order=1)
    bx_y = diff_axis(bx, dy, 0, edge_order=1)
    by_x = diff_axis(by, dx, 1, edge_order=1)
    by_y = diff_axis(by, dy, 0, edge_order=1)
    kxx, kyy = bx_x, by_y
    kxy = 0.5 * (bx_y + by_x)
    tr = kxx + kyy
    kk = kxx * kxx + kyy * kyy + 2.0 * kxy * kxy
    rho_adm = (tr * tr - kk) / (16.0 * math.pi)
    return {"x": x, "y": y, "dx": dx, "dy": dy, "phi": phi,
            "beta_x": bx, "beta_y": by, "rho_adm": rho_adm}


def integrate_signed(field, dx, dy):
    """adm.py:91-97: positive part, *magnitude* of negative part, net."""
    dA = dx * dy
    pos = float(np.sum(field[field > 0.0]) * dA)
    neg = float(np.sum(-field[field < 0.0]) * dA)
    return pos, neg, pos - neg


# --------------------------------------------------------------------------- #
# host scalar algebra (reproduce_rodal_exact.py:51-54, 64-73)
# --------------------------------------------------------------------------- #
def two_point_tail(e_r1, e_r2, r1, r2):
    return float(e_r2 + (r1 / (r2 - r1)) * (e_r2 - e_r1))


def radial_profile_3d(*, rho, sigma, v, extent, n, n_bins=50, dtype="float64", i_begin=0, i_end=None):
    """Radial-shell profile of the 3D rho_adm: the binning of compute_radial_average_z0
    (/root/reference/src/irrotational_warp/tail.py:59-78 -- ``edges = linspace(0, max(R), n_bins + 1)``, masks
    ``(R >= edges[k]) & (R < edges[k+1])``, ``np.mean`` of the selected points, 0 for an empty bin) applied to the
    field of :func:`energy_3d`.  The edges always come from the FULL grid; ``i_begin``/``i_end``
    restrict which planes are binned.  Also returns per-bin sum, count and sum of |rho| (the scale for tolerances)."""
    i_end = n if i_end is None else i_end
    full = energy_3d(rho=rho, sigma=sigma, v=v, extent=extent, n=n, dtype=dtype, i_begin=i_begin, i_end=i_end,
                     return_density=True)
    field = full["e_density"] / full["dV"]
    x, _ = grid_axis(-extent, extent, n, dtype)
    X, Y, Z = np.meshgrid(x[i_begin:i_end], x, x, indexing="ij")
    R = np.sqrt(X * X + Y * Y + Z * Z)
    xm = np.max(x * x)
    r_max = np.sqrt((xm + xm) + xm)
    if i_begin == 0 and i_end == n:
        assert r_max == np.max(R)
    edges = np.linspace(0, r_max, n_bins + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    avg = np.zeros(n_bins)
    tot = np.zeros(n_bins)
    tot_abs = np.zeros(n_bins)
    cnt = np.zeros(n_bins, dtype=np.int64)
    for k in range(n_bins):
        sel = (R >= edges[k]) & (R < edges[k + 1])
        cnt[k] = int(np.sum(sel))
        if cnt[k] > 0:
            avg[k] = np.mean(field[sel])
            tot[k] = np.sum(field[sel])
            tot_abs[k] = np.sum(np.abs(field[sel]))
    return {"edges": edges, "r_bins": centres, "rho_avg": avg, "bin_sum": tot, "bin_abs": tot_abs, "bin_count": cnt,
            "e_pos": full["e_pos"], "e_neg": full["e_neg"]}


# --------------------------------------------------------------------------- #
# Einstein-tensor eigenvalue diagnostic on the z = 0 slice
# --------------------------------------------------------------------------- #
def einstein_eigenvalues(beta_x, beta_y, dx, dy):
    """Restatement of compute_einstein_eigenvalues (/root/reference/src/irrotational_warp/einstein.py:256-309):
    metric (:41-71), pointwise inverse (:74-93), Christoffel symbols from edge_order=1 differences (:96-164),
    Ricci tensor and scalar (:167-226), mixed Einstein tensor (:229-253), eigenvalues, Type-I mask, max real part.
    Array code over the grid instead of the reference's Python loops over points; the accumulation order of every
    sum is the reference's, so everything up to the LAPACK calls is bit-identical (checked by make_golden.py)."""
    beta_x = np.asarray(beta_x, dtype=np.float64)
    beta_y = np.asarray(beta_y, dtype=np.float64)
    ny, nx = beta_x.shape
    g = np.zeros((4, 4, ny, nx))
    g[0, 0] = -1.0 + beta_x ** 2 + beta_y ** 2
    g[0, 1] = g[1, 0] = beta_x
    g[0, 2] = g[2, 0] = beta_y
    g[1, 1] = g[2, 2] = g[3, 3] = 1.0
    # one LAPACK inverse per grid point (gufunc over the leading axes == the reference's per-point calls)
    gi = np.moveaxis(np.linalg.inv(np.moveaxis(g, (0, 1), (2, 3))), (2, 3), (0, 1))

    def d_dx(a):
        return diff_axis(a, dx, 1, edge_order=1)

    def d_dy(a):
        return diff_axis(a, dy, 0, edge_order=1)

    zero = 0.0
    dg = {1: np.stack([np.stack([d_dx(g[a, b]) for b in range(4)]) for a in range(4)]),
          2: np.stack([np.stack([d_dy(g[a, b]) for b in range(4)]) for a in range(4)])}

    def dgc(direction, a, b):          # d_direction g_ab; the slice is stationary and z-independent
        return dg[direction][a, b] if direction in dg else zero

    gamma = np.zeros((4, 4, 4, ny, nx))
    for mu in range(4):
        for nu in range(4):
            for al in range(4):
                for sg in range(4):
                    term = dgc(mu, sg, nu) + dgc(nu, sg, mu) - dgc(sg, mu, nu)
                    gamma[al, mu, nu] += 0.5 * gi[al, sg] * term

    ric = np.zeros((4, 4, ny, nx))
    for mu in range(4):
        for nu in range(4):
            t1 = 0.0
            t1 = t1 + d_dx(gamma[1, mu, nu])
            t1 = t1 + d_dy(gamma[2, mu, nu])
            t2 = 0.0
            if nu in (1, 2):
                dd = d_dx if nu == 1 else d_dy
                for sg in range(4):
                    t2 = t2 + dd(gamma[sg, mu, sg])
            quad = 0.0
            for sg in range(4):
                for rh in range(4):
                    quad = quad + (gamma[sg, rh, sg] * gamma[rh, mu, nu] - gamma[sg, rh, nu] * gamma[rh, mu, sg])
            ric[mu, nu] = t1 - t2 + quad
    scalar = np.zeros((ny, nx))
    for mu in range(4):
        for nu in range(4):
            scalar += gi[mu, nu] * ric[mu, nu]

    mixed = np.zeros((4, 4, ny, nx))
    for mu in range(4):
        for nu in range(4):
            for sg in range(4):
                mixed[mu, nu] += gi[mu, sg] * ric[sg, nu]
    G = np.zeros_like(mixed)
    for mu in range(4):
        for nu in range(4):
            G[mu, nu] = mixed[mu, nu] - 0.5 * (1.0 if mu == nu else 0.0) * scalar

    eig = np.linalg.eigvals(np.moveaxis(G, (0, 1), (2, 3)))          # [ny, nx, 4]
    eig_all = np.moveaxis(eig, 2, 0).astype(complex)
    return {"eig_all": eig_all, "eig_max": np.max(eig_all.real, axis=0), "ricci_scalar": scalar,
            "is_type_i": np.all(np.abs(eig_all.imag) < 1e-8, axis=0), "G_mixed": G, "gamma": gamma, "g_con": gi}

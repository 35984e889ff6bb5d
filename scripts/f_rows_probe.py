#!/usr/bin/env python3
"""Timings of the SURVEY section-8(f) rows on the GPU box (development aid):
the 3D radial profile (f3) at n = 512 / 1024 and the Einstein diagnostic (f4) at n = 101 / 301 / 1001, with the
oracle (CPU, NumPy + LAPACK) beside the latter at n = 101."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200 import tail
from irrotational_warp_lab_b200.einstein import compute_einstein_eigenvalues
from irrotational_warp_lab_b200.slice2d import compute_slice_z0

for n in (256, 512, 1024):
    for rep in range(2):
        t0 = time.perf_counter()
        p = tail.radial_profile_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n)
        dt = time.perf_counter() - t0
    print(f"radial_profile_3d n={n} ({p.get('path')}): kernels {p['kernel_ms']:.2f} ms, call {dt*1e3:.2f} ms, "
          f"{n**3/p['kernel_ms']/1e6:.1f} Gpts/s; points binned {int(p['bin_count'].sum())} of {n**3}; "
          f"e_net {p['energies']['e_net']:.6e}", flush=True)
t = tail.compute_tail_correction_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=512)
print(f"tail fit 3D n=512: exponent {t.exponent:.3f} amplitude {t.amplitude:.3e} rms {t.fit_residual_rms:.3f} "
      f"tail+ {t.tail_integral_pos:.3e} tail- {t.tail_integral_neg:.3e}", flush=True)

for n in (101, 301, 1001):
    s = compute_slice_z0(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=n)
    for rep in range(2):
        t0 = time.perf_counter()
        r = compute_einstein_eigenvalues(s.beta_x, s.beta_y, dx=s.dx, dy=s.dy)
        dt = time.perf_counter() - t0
    print(f"einstein n={n}: call {dt*1e3:.2f} ms; type-I fraction {r.is_type_i.mean():.4f}; max|eig_max| {np.abs(r.eig_max).max():.4e}", flush=True)
    if n == 101:
        from oracle import oracle_np
        t0 = time.perf_counter()
        o = oracle_np.einstein_eigenvalues(s.beta_x, s.beta_y, s.dx, s.dy)
        print(f"   oracle (NumPy + LAPACK, CPU) n=101: {(time.perf_counter()-t0)*1e3:.1f} ms; type-I mismatches "
              f"{int((o['is_type_i'] != r.is_type_i).sum())}", flush=True)

#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel variant once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200 import backend, slice2d
for n, dtype in ((45, "float64"), (33, "float32")):
    r = backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n, dtype=dtype, emit_rho=(n == 45))
    print(n, dtype, r["energies"]["e_pos"], r["counts"]["neg"])
r = backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40, shells=[5.0, 10.0, 20.0, 40.0])
print("4 shells", r["counts"]["shell_points"])
a = backend.compute_energy_axisymmetric(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=130, ny=70)
print("axisym", a["energies"]["e_pos"])
s = slice2d.compute_slice_z0(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=41)
print("slice", slice2d.integrate_signed(s, s.dx, s.dy))
from irrotational_warp_lab_b200 import tail
p = tail.radial_profile_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=45, n_bins=30)
print("radial profile (fused kernel)", int(p["bin_count"].sum()))

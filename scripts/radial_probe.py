#!/usr/bin/env python3
"""One radial profile at n (development aid; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200 import tail
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for rep in range(2):
    p = tail.radial_profile_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n)
print(f"n={n}: kernels {p['kernel_ms']:.2f} ms, {n**3/p['kernel_ms']/1e6:.1f} Gpoints/s")

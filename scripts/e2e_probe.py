#!/usr/bin/env python3
"""Where the host time of one public-API call goes (development aid): wall time of backend.compute_energy_3d against the
CUDA-event time of its kernels, and the pieces of the call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200 import backend
from irrotational_warp_lab_b200.engine import get_engine, Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = get_engine()
for _ in range(3):
    backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n, shells=[40.0, 60.0])
ts, ks, tot = [], [], []
for _ in range(10):
    t0 = time.perf_counter()
    r = backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n, shells=[40.0, 60.0])
    ts.append((time.perf_counter() - t0) * 1e3); ks.append(r["engine"]["kernel_ms"])
print(f"n={n}: API wall {np.median(ts):.3f} ms, kernel events {np.median(ks):.3f} ms, overhead {np.median(ts) - np.median(ks):.3f} ms")
x = np.linspace(-66.0, 66.0, n); dx = float(x[1] - x[0])
t0 = time.perf_counter()
for _ in range(100):
    x = np.linspace(-66.0, 66.0, n); dx = float(x[1] - x[0])
print(f"  linspace: {(time.perf_counter() - t0) * 10:.4f} ms per call (x3 per evaluation)")
t0 = time.perf_counter()
for _ in range(100):
    p, keep = Engine.make_params3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0])
print(f"  make_params3d: {(time.perf_counter() - t0) * 10:.4f} ms")
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    r = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0])
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"  eng.energy3d wall {np.median(ts):.3f} ms, kernel {r['kernel_ms']:.3f} ms, C total_ms {r['total_ms']:.3f}")

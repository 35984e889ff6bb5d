"""Kernel-event time of the axisymmetric evaluator (k_axisym + k_reduce_tiles + the row read-back) for a few sizes.
usage: python scripts/axisym_probe.py [reps]   (development aid; honours IWB_LIB with IWB_ALLOW_LIB_OVERRIDE=1)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200 import backend  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for dtype in ("float64", "float32"):
    for nx, ny in ((2400, 1200), (1200, 600), (4800, 2400)):
        kw = dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=nx, ny=ny, dtype=dtype)
        for _ in range(3):
            run = backend.compute_energy_axisymmetric(**kw)
        ks, ws = [], []
        for _ in range(reps):
            t0 = time.perf_counter()
            run = backend.compute_energy_axisymmetric(**kw)
            ws.append(time.perf_counter() - t0)
            ks.append(run["engine"]["kernel_ms"])
        pts = nx * ny
        print(f"axisym {dtype} {nx}x{ny}: kernel-event min {min(ks):.4f} ms median {sorted(ks)[len(ks)//2]:.4f} ms "
              f"({pts / min(ks) / 1e6:.1f} Gpoints/s), API call min {min(ws)*1e3:.3f} ms; "
              f"e_pos {run['energies']['e_pos']!r} e_neg {run['energies']['e_neg']!r}", flush=True)

#!/usr/bin/env python3
"""Time the fused 3D kernel at one size (development aid).  usage: k3_probe.py n [reps] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200.engine import get_engine

n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dtype = sys.argv[3] if len(sys.argv) > 3 else "float64"
mode = os.environ.get("IWB_MODE", "exact")
eng = get_engine()
x = np.linspace(-66.0, 66.0, n, dtype=np.float32 if dtype.startswith("float32") else np.float64); dx = float(x[1] - x[0])
ts = []
for rep in range(reps):
    r = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype=dtype, shells=[40.0, 60.0], mode=mode)
    ts.append(r["kernel_ms"])
best = min(ts[1:]) if reps > 1 else ts[0]
print(f"mode={mode} cfg={os.environ.get('IWB_K3_CONFIG','default')} static_pct={os.environ.get('IWB_K3_STATIC_PCT','auto')} n={n} {dtype} best {best:.3f} ms "
      f"{n**3/best/1e6:.2f} Gpts/s all={['%.2f'%t for t in ts]} e_pos={r['e_pos']!r} neg={r['cnt_neg']}", flush=True)

if len(sys.argv) > 4:      # slab probe: planes [a, b)
    a, b = int(sys.argv[4]), int(sys.argv[5])
    ts = []
    for rep in range(reps):
        r = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype=dtype, shells=[40.0, 60.0], i_begin=a, i_end=b)
        ts.append(r["kernel_ms"])
    best = min(ts[1:])
    print(f"  slab [{a},{b}) best {best:.3f} ms  {(b-a)*n*n/best/1e6:.2f} Gpts/s-equivalent per GPU", flush=True)

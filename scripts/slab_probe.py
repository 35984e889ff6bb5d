#!/usr/bin/env python3
"""Time the fused kernel on the slabs of a P-way decomposition of n=1024^3 (development aid): slab_probe.py P [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200.engine import get_engine
P = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1024
eng = get_engine()
x = np.linspace(-66.0, 66.0, n); dx = float(x[1] - x[0])
out = []
for r in range(P):
    a, b = r * n // P, (r + 1) * n // P
    ts = []
    for rep in range(reps):
        res = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0], i_begin=a, i_end=b)
        ts.append(res["kernel_ms"])
    out.append(min(ts[1:]))
print(f"static_pct={os.environ.get("IWB_K3_STATIC_PCT","auto")} P={P} slab ms: " + " ".join(f"{t:.3f}" for t in out) + f"  max {max(out):.3f} mean {sum(out)/len(out):.3f}", flush=True)

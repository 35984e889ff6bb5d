#!/usr/bin/env python3
"""Error and speed of the float32 variants per grid size (GPU box; development aid / evidence for DESIGN.md section 2).

float32 = IWB_F32_REF: the reference's --dtype float32 as NumPy >= 2 evaluates it (bit-identical to it on the goldens).
The north star's bar for float32 is 1e-5 relative against the reference's float32 path; float64 is shown for scale.
(profiles/r2_f32_error_table.log also holds the rows of the all-float32 phi variant that was measured with this script --
slower, and 8e-4 off at n = 1024 -- and then removed from the library)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200.engine import get_engine

eng = get_engine()
rel = lambda a, b: abs(a - b) / abs(b)
print(f"{'n':>5} {'dtype':>15} {'e_pos':>22} {'e_neg':>23} {'rel vs f32-ref (E+, E-)':>28} {'kernel ms':>10} {'Gpts/s':>8}")
for n in (100, 256, 512, 1024):
    out = {}
    for dtype in ("float64", "float32"):
        x = np.linspace(-66.0, 66.0, n, dtype=np.float32 if dtype.startswith("float32") else np.float64)
        dx = float(x[1] - x[0])
        best = None
        for rep in range(4):
            r = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype=dtype, shells=[40.0, 60.0])
            best = r["kernel_ms"] if best is None or (rep > 0 and r["kernel_ms"] < best) else best
        out[dtype] = (r, best)
    ref = out["float32"][0]
    for dtype in ("float64", "float32"):
        r, ms = out[dtype]
        print(f"{n:>5} {dtype:>15} {r['e_pos']!r:>22} {r['e_neg']!r:>23} "
              f"{rel(r['e_pos'], ref['e_pos']):>13.2e} {rel(r['e_neg'], ref['e_neg']):>13.2e} {ms:>10.3f} {n ** 3 / ms / 1e6:>8.1f}", flush=True)

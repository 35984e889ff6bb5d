"""Throughput of iwb_energy3d_batch for grids of a few sizes.  usage: python scripts/batch_probe.py [npoints] [reps]
(development aid; the IWB_BATCH_SPLIT switch it was written for -- SM partitions per concurrent evaluation -- was measured
with it and removed, profiles/r2_ab18_batch_split.log)"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200 import optimize  # noqa: E402
from irrotational_warp_lab_b200.engine import get_engine  # noqa: E402

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 50
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = get_engine(0)
rng = np.random.default_rng(7)
params = [(float(s), float(v)) for s, v in zip(rng.uniform(2.0, 8.0, npts), rng.uniform(0.8, 2.0, npts))]
for n in (100, 256, 400, 512):
    pts = optimize._points_3d(params, 5.0, 66.0, n, "float64")
    for _ in range(2):
        res = eng.energy3d_batch(pts)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        res = eng.energy3d_batch(pts)
        best = min(best, time.perf_counter() - t0)
    digest = hashlib.sha256(repr([(r["e_pos"], r["e_neg"], r["cnt_neg"], r["cnt_pos"]) for r in res]).encode()).hexdigest()[:16]
    print(f"batch of {npts} x n={n}^3: {best*1e3:.3f} ms = "
          f"{npts/best:.0f} evaluations/s = {npts*n**3/best/1e9:.1f} Gpoints/s; results sha {digest}", flush=True)

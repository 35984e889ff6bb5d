#!/usr/bin/env python3
"""Determinism soak: repeated evaluations must be bit-identical (a data race in the ring / split-barrier
pipeline would show up as run-to-run differences).  usage: soak.py [reps]"""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200.engine import get_engine
from oracle import oracle_np

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
eng = get_engine()
bad = 0
for n, dtype, mode in [(37, "float64", "exact"), (64, "float64", "exact"), (100, "float32", "exact"), (129, "float64", "exact"),
                       (200, "float64", "exact"), (161, "float64", "fast"), (320, "float64", "exact")]:
    x, dx = oracle_np.grid_axis(-66.0, 66.0, n, dtype)
    kw = dict(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype=dtype, shells=[40.0, 60.0], mode=mode)
    ref = None
    for rep in range(reps):
        rho = np.empty((n, n, n)) if rep % 4 == 0 else None       # emitting rho must not change the sums
        r = eng.energy3d(rho_out=rho, **kw)
        key = (r["e_pos"], r["e_neg"], r["cnt_pos"], r["cnt_neg"], r["cnt_zero"], tuple(r["shell_pos"]), tuple(r["shell_neg"]), tuple(r["shell_cnt"]))
        h = hashlib.sha256(rho.tobytes()).hexdigest() if rho is not None else None
        if ref is None:
            ref, href = key, h
        else:
            if key != ref or (h is not None and h != href):
                bad += 1
                print("MISMATCH", n, dtype, mode, rep, key, ref)
    print(f"n={n} {dtype} {mode}: {reps} runs identical={bad == 0}", flush=True)
print("SOAK", "OK" if bad == 0 else f"FAILED ({bad})")
sys.exit(1 if bad else 0)

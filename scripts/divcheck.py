#!/usr/bin/env python3
"""Long run of the division / square-root self-test (development aid): divcheck.py log2_count nseeds"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200.engine import get_engine
eng = get_engine()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tot = 0
for seed in range(100, 100 + ns):
    t = time.time()
    r = eng.selftest_divsqrt(1 << lg, seed=seed, wide=False)
    tot += 1 << lg
    print(f"lib={os.environ.get('IWB_LIB','default')} seed={seed} pairs=2^{lg} {r} {time.time()-t:.1f}s", flush=True)

#!/usr/bin/env python3
"""Runs the register-resident DFMA / FFMA microbenchmark of the engine (roofline denominator; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200.engine import get_engine
eng = get_engine()
for fp32 in (False, True):
    r = eng.measure_fma_peak(fp32)
    print(("fp32" if fp32 else "fp64"), r, flush=True)

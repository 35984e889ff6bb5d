#!/usr/bin/env python3
"""Time the fused kernel on the interleaved tile shards of a P-way run of n=1024^3, one after the other on one GPU
(development aid): shard_probe.py P [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irrotational_warp_lab_b200.engine import get_engine
from irrotational_warp_lab_b200._capi import ROW_WIDTH, TILE_CHAINS
P = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1024
eng = get_engine()
x = np.linspace(-66.0, 66.0, n); dx = float(x[1] - x[0])
st = torch.cuda.current_stream()
buf = torch.zeros((64, TILE_CHAINS, ROW_WIDTH), dtype=torch.float64, device="cuda:0")
out = []
for r in range(P):
    plan = eng.plan3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0],
                      tile_stride=P, tile_first=r)
    ts = []
    for rep in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); plan.run_async(buf.data_ptr(), st.cuda_stream); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out.append(min(ts[1:])); plan.close()
print(f"static_pct={os.environ.get('IWB_K3_STATIC_PCT','auto')} tile shards P={P} ms: " + " ".join(f"{t:.3f}" for t in out) +
      f"  max {max(out):.3f} mean {sum(out)/len(out):.3f} ideal {10.14/P:.3f}", flush=True)

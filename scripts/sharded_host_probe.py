#!/usr/bin/env python3
"""Host cost of one sharded public-API step, piece by piece, on ONE GPU with a one-rank peer group (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from irrotational_warp_lab_b200._capi import ROW_WIDTH
from irrotational_warp_lab_b200.backend import _grid
from irrotational_warp_lab_b200.engine import Engine, PeerGroup, get_engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = get_engine()
pg = PeerGroup(eng, 0, 1, 256, lambda blob: [blob])
dev = torch.device("cuda", 0)
x, dx = _grid(-66.0, 66.0, n, "float64")
kw = dict(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0], eps=1e-12)
def timeit(f, reps=200):
    f(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e6
print(f"n={n}")
print(f"  _grid (linspace)            {timeit(lambda: _grid(-66.0, 66.0, n, 'float64')):8.1f} us")
print(f"  torch.empty(row)            {timeit(lambda: torch.empty(ROW_WIDTH, dtype=torch.float64, device=dev)):8.1f} us")
print(f"  current_stream              {timeit(lambda: torch.cuda.current_stream(dev).cuda_stream):8.1f} us")
print(f"  make_params3d               {timeit(lambda: Engine.make_params3d(tile_stride=1, tile_first=0, **kw)):8.1f} us")
row = torch.empty(ROW_WIDTH, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
def enqueue():
    pg.energy3d_tiles_exchange_async(row.data_ptr(), st, **kw)
torch.cuda.synchronize()
t = []
for _ in range(50):
    torch.cuda.synchronize(); t0 = time.perf_counter(); enqueue(); t.append((time.perf_counter() - t0) * 1e6); torch.cuda.synchronize()
print(f"  exchange_async (enqueue)    {np.median(t):8.1f} us   (make_params3d + C staging + 3 launches)")
t = []
for _ in range(50):
    enqueue(); torch.cuda.synchronize(); t0 = time.perf_counter(); eng.read_row(row.data_ptr(), 2, st); t.append((time.perf_counter() - t0) * 1e6)
print(f"  read_row on an idle stream  {np.median(t):8.1f} us")
t = []
for _ in range(50):
    torch.cuda.synchronize(); t0 = time.perf_counter(); enqueue(); r = eng.read_row(row.data_ptr(), 2, st); t.append((time.perf_counter() - t0) * 1e6)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); enqueue(); b.record(); torch.cuda.synchronize()
print(f"  enqueue + read_row (blocking step) {np.median(t):8.1f} us, of which device {a.elapsed_time(b) * 1e3:.1f} us")

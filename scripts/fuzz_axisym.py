"""Seeded random configurations of the axisymmetric evaluator against the C oracle (development aid; GPU box).
usage: python scripts/fuzz_axisym.py [ncases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irrotational_warp_lab_b200 import backend  # noqa: E402
from oracle import oracle_c  # noqa: E402

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2027)
bad = 0
for c in range(ncases):
    rho = float(np.round(rng.uniform(0.5, 8.0), 3))
    sigma = float(np.round(rng.uniform(0.6, 12.0), 3))
    v = float(np.round(rng.uniform(-3.0, 3.0), 3))
    extent = float(np.round(rho * rng.uniform(1.5, 14.0), 3))
    nx, ny = int(rng.integers(5, 400)), int(rng.integers(3, 200))
    dtype = "float32" if c % 5 == 4 else "float64"
    kw = dict(rho=rho, sigma=sigma, v=v, extent=extent, nx=nx, ny=ny)
    shells = [2.0 * rho, 0.5 * extent]
    run = backend.compute_energy_axisymmetric(dtype=dtype, shells=shells, emit_rho=True, **kw)
    o = oracle_c.energy_axisym(dtype=dtype, shells=shells, return_density=True, **kw)
    scale = float(np.abs(o["rho_adm"]).max()) or 1.0
    drho = float(np.abs(run["rho_adm"] - o["rho_adm"]).max()) / scale
    ndiff = int(np.count_nonzero(run["rho_adm"] != o["rho_adm"]))
    counts_ok = (run["counts"]["neg"], run["counts"]["pos"], run["counts"]["zero"]) == (o["cnt_neg"], o["cnt_pos"], o["cnt_zero"])
    shells_ok = [run["counts"]["shell_points"][r] for r in shells] == [s["cnt"] for s in o["shells"]]
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)      # noqa: E731
    e_err = max(rel(run["energies"]["e_pos"], o["e_pos"]), rel(run["energies"]["e_neg"], o["e_neg"]))
    tol_rho, tol_e = (1e-10, 1e-10) if dtype == "float64" else (1e-4, 1e-5)      # axisymmetric pointwise bound: DESIGN.md section 2
    ok = counts_ok and shells_ok and drho <= tol_rho and e_err <= tol_e
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} case {c:3d} {dtype} nx={nx:3d} ny={ny:3d} rho={rho} sigma={sigma} v={v} extent={extent}: "
          f"points differing {ndiff}, max|d rho|/max|rho| {drho:.2e}, energies {e_err:.2e}, counts {counts_ok} shells {shells_ok}", flush=True)
print(f"{bad} of {ncases} cases outside the bars")

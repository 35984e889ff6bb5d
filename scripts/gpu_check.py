#!/usr/bin/env python3
"""Quick GPU parity + timing probe (development aid; the real tests live in tests/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irrotational_warp_lab_b200 import backend
from irrotational_warp_lab_b200.engine import get_engine
from oracle import oracle_c

eng = get_engine()
print("device:", eng.name, eng.sm_count, eng.cc, flush=True)
g = json.load(open("tests/golden/golden_3d.json"))["cases"]
ok_all = True
for name, rec in sorted(g.items()):
    kw = rec["params"]
    shells = [s["r"] for s in rec["shells"]]
    run = backend.compute_energy_3d(shells=shells, emit_rho=True, **kw)
    en = run["energies"]
    o = oracle_c.energy_3d(shells=shells, return_density=True, **kw)
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)
    rp, rn = rel(en["e_pos"], rec["energies"]["e_pos"]), rel(en["e_neg"], rec["energies"]["e_neg"])
    cnt_ok = (run["counts"]["neg"] == rec["cnt_neg"] and run["counts"]["pos"] == rec["cnt_pos"]
              and run["counts"]["zero"] == rec["cnt_zero"])
    sh_ok = all(run["counts"]["shell_points"][s["r"]] == s["cnt"] for s in rec["shells"])
    sh_rel = max([rel(run["_e_at_r"](s["r"])[0], s["e_pos"]) for s in rec["shells"]] +
                 [rel(run["_e_at_r"](s["r"])[1], s["e_neg"]) for s in rec["shells"]])
    diff = np.abs(run["rho_adm"] - o["rho_adm"])
    nbad = int(np.count_nonzero(run["rho_adm"] != o["rho_adm"]))
    print(f"{name:18s} rel_pos {rp:.2e} rel_neg {rn:.2e} shells {sh_rel:.2e} counts {cnt_ok} shellcnt {sh_ok} "
          f"rho!=oracle {nbad}/{diff.size} maxabs {diff.max():.3e} (max|rho| {np.abs(o['rho_adm']).max():.3e}) "
          f"kernel {run['engine']['kernel_ms']:.3f} ms", flush=True)
    ok_all &= cnt_ok and sh_ok and rp < 1e-10 and rn < 1e-10
print("ALL OK" if ok_all else "MISMATCH", flush=True)

print("fp64 FMA peak:", eng.measure_fma_peak(False), "fp32:", eng.measure_fma_peak(True), flush=True)
for d in (2.0 * 0.12903225806451513, 16.0 * np.pi, 2.6666666666666572):
    print("constdiv mismatches for divisor", d, ":", eng.selftest_constdiv(d, 1 << 28, 7), flush=True)

for n in (256, 512, 1024):
    x = np.linspace(-66.0, 66.0, n); dx = float(x[1] - x[0])
    best = 1e30
    for rep in range(4):
        r = eng.energy3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=5.0, sigma=4.0, v=1.0, dtype="float64", shells=[40.0, 60.0])
        best = min(best, r["kernel_ms"])
    print(f"n={n} kernel {best:.3f} ms  {n**3/best/1e6:.2f} Gpts/s  total_ms {r['total_ms']:.3f} e_pos {r['e_pos']!r} e_neg {r['e_neg']!r} neg {r['cnt_neg']}", flush=True)

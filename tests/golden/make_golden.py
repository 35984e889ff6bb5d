#!/usr/bin/env python3
"""Freeze golden vectors from the UNMODIFIED reference (build container only).

Run from the repo root:

    python tests/golden/make_golden.py            # writes tests/golden/*.json
    python tests/golden/make_golden.py --check-oracle   # also diffs oracle/ pointwise

It imports ``/root/reference`` in place (``src`` for the package, ``scripts`` for
``reproduce_rodal_exact``), wraps ``_split_pos_neg`` to capture the e-density
array and stores, per case: energies, shell sums, sign counts,
shell point counts, a SHA-256 of the e-density bytes and ~200 sampled point
values.  Nothing under ``tests/`` reads ``/root/reference`` at test time.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IWB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(REF, "scripts"))
sys.path.insert(0, ROOT)

import reproduce_rodal_exact as ref  # noqa: E402
from irrotational_warp import adm as ref_adm  # noqa: E402


def _capture(fn, **kw):
    box = {}
    orig = ref._split_pos_neg

    def spy(xp, e_density):
        box["e"] = np.array(e_density, copy=True)
        return orig(xp, e_density)

    ref._split_pos_neg = spy
    try:
        run = fn(**kw)
    finally:
        ref._split_pos_neg = orig
    return run, box["e"]


def _samples(e, count=200, seed=1234):
    rng = np.random.default_rng(seed)
    flat = e.reshape(-1)
    idx = np.unique(np.concatenate([
        rng.integers(0, flat.size, size=count),
        np.array([0, 1, 2, flat.size - 1, flat.size - 2, flat.size // 2]),
        np.argsort(np.abs(flat))[:16],           # the noise core (H1)
        np.argsort(-np.abs(flat))[:8],
    ]))
    return {"index": idx.tolist(), "value_hex": [float(flat[i]).hex() for i in idx]}


def _record(run, e, shells):
    e_at_r = run["_e_at_r"]
    rec = {
        "grid": run["grid"],
        "dtype": run["dtype"],
        "energies": run["energies"],
        "cnt_pos": int(np.count_nonzero(e > 0.0)),
        "cnt_neg": int(np.count_nonzero(e < 0.0)),
        "cnt_zero": int(np.count_nonzero(e == 0.0)),
        "abs_sum": float(np.sum(np.abs(e))),
        "e_sha256": hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest(),
        "e_dtype": str(e.dtype),
        "e_max_abs": float(np.max(np.abs(e))),
        "shells": [],
        "samples": _samples(e),
    }
    for r_lim in shells:
        p, q = e_at_r(r_lim)
        rec["shells"].append({"r": r_lim, "e_pos": p, "e_neg": q})
    return rec


def shell_counts_3d(n, extent, dtype, shells):
    dt = np.float32 if dtype == "float32" else np.float64
    x = np.linspace(-extent, extent, n, dtype=dt)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    return [int(np.count_nonzero(r <= s)) for s in shells]


CASES_3D = [
    # name, params
    ("n40", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40, dtype="float64")),
    ("n60", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=60, dtype="float64")),
    ("n80", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=80, dtype="float64")),
    ("n100", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=100, dtype="float64")),
    ("n101", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=101, dtype="float64")),
    ("n128", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=128, dtype="float64")),
    ("n200", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=200, dtype="float64")),
    ("n256", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=256, dtype="float64")),
    ("n100_f32", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=100, dtype="float32")),
    ("n256_f32", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=256, dtype="float32")),
    ("n80_v3", dict(rho=5.0, sigma=4.0, v=3.0, extent=66.0, n=80, dtype="float64")),
    ("n64_rho10_sig3", dict(rho=10.0, sigma=3.0, v=1.5, extent=132.0, n=64, dtype="float64")),
    ("n48_thin", dict(rho=2.0, sigma=8.0, v=0.8, extent=12.0, n=48, dtype="float64")),
    ("n33_small_extent", dict(rho=5.0, sigma=4.0, v=1.0, extent=8.0, n=33, dtype="float64")),
    ("n7_tiny", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=7, dtype="float64")),
    ("n3_min", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=3, dtype="float64")),
    ("n50_v0", dict(rho=5.0, sigma=4.0, v=0.0, extent=66.0, n=50, dtype="float64")),
]

CASES_AXI = [
    ("axi_2400x1200", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=2400, ny=1200, dtype="float64")),
    ("axi_1200x600", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=1200, ny=600, dtype="float64")),
    ("axi_1200x600_v3", dict(rho=5.0, sigma=4.0, v=3.0, extent=66.0, nx=1200, ny=600, dtype="float64")),
    ("axi_301x77", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=301, ny=77, dtype="float64")),
    ("axi_300x150_f32", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=300, ny=150, dtype="float32")),
    ("axi_5x3", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=5, ny=3, dtype="float64")),
]

CASES_SLICE = [
    ("slice_n101", dict(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=101)),
    ("slice_n71", dict(rho=10.0, sigma=5.0, v=1.2, extent=20.0, n=71)),
    ("slice_n64", dict(rho=10.0, sigma=2.0, v=0.8, extent=30.0, n=64)),
    ("slice_n2", dict(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=2)),
    ("slice_v0", dict(rho=10.0, sigma=3.0, v=0.0, extent=20.0, n=33)),
]


CASES_EINSTEIN = [
    ("ein_plot_n41", dict(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=41)),
    ("ein_plot_n101", dict(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=101)),
    ("ein_subluminal_n61", dict(rho=5.0, sigma=4.0, v=0.5, extent=12.0, n=61)),
    ("ein_minkowski_n21", dict(field="minkowski")),
    ("ein_sine_n21", dict(field="sine")),
]


def einstein_test_field(kind):
    """The two inputs of the reference's tests/test_einstein.py (flat space; a small harmonic shift)."""
    n = 21
    if kind == "minkowski":
        return np.zeros((n, n)), np.zeros((n, n)), 0.5, 0.5
    x = np.linspace(-5, 5, n)
    X, _ = np.meshgrid(x, x)
    d = float(x[1] - x[0])
    bx = 0.01 * np.sin(2 * np.pi * X / 10.0)
    return bx, np.zeros_like(bx), d, d


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-oracle", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args(argv)

    out3d, outaxi, outslice = {}, {}, {}
    for name, kw in CASES_3D:
        if args.only and args.only not in name:
            continue
        shells = [8.0 * kw["rho"], 12.0 * kw["rho"], 1.7 * kw["rho"]]
        run, e = _capture(ref.compute_energy_3d, backend="numpy", **kw)
        rec = _record(run, e, shells)
        rec["params"] = kw
        for s, c in zip(rec["shells"], shell_counts_3d(kw["n"], kw["extent"], kw["dtype"], shells)):
            s["cnt"] = c
        out3d[name] = rec
        print(name, rec["energies"]["e_pos"], rec["energies"]["e_neg"], rec["cnt_neg"], rec["cnt_pos"])
        if args.check_oracle:
            from oracle import oracle_np
            o = oracle_np.energy_3d(shells=shells, return_density=True, **kw)
            same = np.array_equal(o["e_density"], e, equal_nan=True)
            print("   oracle_np pointwise bit-equal:", same,
                  "| e_pos equal:", o["e_pos"] == rec["energies"]["e_pos"],
                  "| e_neg equal:", o["e_neg"] == rec["energies"]["e_neg"])
            assert same
            if kw["n"] >= 40:
                oc = oracle_np.energy_3d_chunked(shells=shells, chunk=16, **kw)
                print("   chunked rel diff:", abs(oc["e_pos"] - o["e_pos"]) / max(abs(o["e_pos"]), 1e-300),
                      "counts equal:", oc["cnt_neg"] == o["cnt_neg"] and oc["cnt_pos"] == o["cnt_pos"])
                assert oc["cnt_neg"] == o["cnt_neg"] and oc["cnt_pos"] == o["cnt_pos"]

    for name, kw in CASES_AXI:
        if args.only and args.only not in name:
            continue
        shells = [8.0 * kw["rho"], 12.0 * kw["rho"]]
        run, e = _capture(ref.compute_energy_axisymmetric, backend="numpy", **kw)
        rec = _record(run, e, shells)
        rec["params"] = kw
        outaxi[name] = rec
        print(name, rec["energies"]["e_pos"], rec["energies"]["e_neg"])
        if args.check_oracle:
            from oracle import oracle_np
            o = oracle_np.energy_axisym(shells=shells, return_density=True, **kw)
            same = np.array_equal(o["e_density"], e, equal_nan=True)
            print("   oracle_np pointwise bit-equal:", same)
            assert same

    for name, kw in CASES_SLICE:
        if args.only and args.only not in name:
            continue
        res = ref_adm.compute_slice_z0(**kw)
        pos, neg, net = ref_adm.integrate_signed(res.rho_adm, res.dx, res.dy)
        rec = {
            "params": kw, "dx": res.dx, "dy": res.dy,
            "e_pos": pos, "e_neg": neg, "e_net": net,
            "cnt_pos": int(np.count_nonzero(res.rho_adm > 0)),
            "cnt_neg": int(np.count_nonzero(res.rho_adm < 0)),
            "cnt_zero": int(np.count_nonzero(res.rho_adm == 0)),
            "rho_min": float(res.rho_adm.min()), "rho_max": float(res.rho_adm.max()),
        }
        for fld in ("rho_adm", "phi", "beta_x", "beta_y"):
            arr = np.ascontiguousarray(getattr(res, fld))
            rec[fld + "_sha256"] = hashlib.sha256(arr.tobytes()).hexdigest()
            rec[fld + "_samples"] = _samples(arr, count=64)
        outslice[name] = rec
        print(name, pos, neg, net)
        if args.check_oracle:
            from oracle import oracle_np
            o = oracle_np.slice_z0(**kw)
            for fld in ("rho_adm", "phi", "beta_x", "beta_y"):
                assert np.array_equal(o[fld], getattr(res, fld), equal_nan=True), fld
            assert oracle_np.integrate_signed(o["rho_adm"], o["dx"], o["dy"]) == (pos, neg, net)
            print("   oracle_np fields bit-equal: True")

    # Einstein-tensor diagnostic (einstein.py): inputs are the reference's own slice fields and the two fields of
    # the reference's tests/test_einstein.py
    from irrotational_warp import einstein as ref_ein
    outein = {}
    for name, kw in CASES_EINSTEIN:
        if args.only and args.only not in name:
            continue
        if "field" in kw:
            bx, by, dx, dy = einstein_test_field(kw["field"])
        else:
            res = ref_adm.compute_slice_z0(**kw)
            bx, by, dx, dy = res.beta_x, res.beta_y, res.dx, res.dy
        er = ref_ein.compute_einstein_eigenvalues(bx, by, dx=dx, dy=dy)
        eig_sorted = np.sort_complex(np.moveaxis(er.eig_all, 0, -1))           # [ny, nx, 4]
        rec = {"params": kw, "dx": dx, "dy": dy, "shape": list(bx.shape),
               "beta_x_sha256": hashlib.sha256(np.ascontiguousarray(bx).tobytes()).hexdigest(),
               "beta_y_sha256": hashlib.sha256(np.ascontiguousarray(by).tobytes()).hexdigest(),
               "type_i_count": int(er.is_type_i.sum()),
               "ricci_abs_max": float(np.abs(er.ricci_scalar).max()),
               "eig_max_abs_max": float(np.abs(er.eig_max).max()),
               "ricci_scalar_samples": _samples(er.ricci_scalar, count=96),
               "eig_max_samples": _samples(er.eig_max, count=96),
               "eig_sorted_re_samples": _samples(np.ascontiguousarray(eig_sorted.real), count=96),
               "eig_sorted_im_samples": _samples(np.ascontiguousarray(eig_sorted.imag), count=96),
               "type_i_samples": _samples(er.is_type_i.astype(np.float64), count=96)}
        outein[name] = rec
        print(name, rec["type_i_count"], rec["ricci_abs_max"], rec["eig_max_abs_max"])
        if args.check_oracle:
            from oracle import oracle_np
            o = oracle_np.einstein_eigenvalues(bx, by, dx, dy)
            for fld in ("eig_all", "eig_max", "ricci_scalar", "is_type_i"):
                assert np.array_equal(o[fld], getattr(er, fld), equal_nan=True), fld
            print("   oracle_np einstein fields bit-equal: True")

    meta = {"numpy": np.__version__, "generator": "tests/golden/make_golden.py",
            "reference": "DawsonInstitute/irrotational-warp-lab (unmodified, NumPy path)"}
    if not args.only:
        for fname, payload in (("golden_3d.json", out3d), ("golden_axisym.json", outaxi),
                               ("golden_slice.json", outslice), ("golden_einstein.json", outein)):
            with open(os.path.join(HERE, fname), "w") as f:
                json.dump({"meta": meta, "cases": payload}, f, indent=1, sort_keys=True)
            print("wrote", fname)


if __name__ == "__main__":
    main()

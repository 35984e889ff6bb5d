"""Command-line front ends with the reference's flags and JSON schemas, running on the engine.

  * ``reproduce``  <- ``scripts/reproduce_rodal_exact.py`` ``main``  (/root/reference/scripts/reproduce_rodal_exact.py:244-341)
  * ``sweep``      <- ``scripts/sweep_superluminal.py`` ``main``     (/root/reference/scripts/sweep_superluminal.py:124-200)
  * ``slice``      <- ``python -m irrotational_warp plot`` minus the image: the summary JSON of ``viz.plot_slice``
                      (/root/reference/src/irrotational_warp/viz.py:30-105, flags of cli.py:16-25), with the optional
                      ``--einstein`` and ``--tail-correction`` blocks

  * ``optimize``   <- ``python -m irrotational_warp optimize``           (/root/reference/src/irrotational_warp/cli.py:56-77, 179-276)
                      plus ``--objective volume3d`` (BASELINE config 5)
  * ``convergence``<- ``scripts/convergence_study_3d.py``              (/root/reference/scripts/convergence_study_3d.py:21-183)

    python -m irrotational_warp_lab_b200.cli reproduce --mode 3d --n 256 --out results/repro.json
    python -m irrotational_warp_lab_b200.cli sweep --mode 3d --n 80 --v-steps 20 --out results/sweep.json

The payloads keep the reference's keys (``meta, params, run, tail_2pt, wall_clock_s`` and
``meta, sweep, total_time_s``), so ``convergence_study_3d.py``, ``plot_superluminal.py``,
``make_paper_figures.py`` and the notebook consume them unchanged; ``run`` gains ``counts`` and ``engine``.
``--backend`` accepts only ``b200`` here (the NumPy / CuPy paths live in the reference).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from datetime import datetime, timezone

from . import backend as b200
from . import sweeps


def _git_info() -> dict:
    try:
        run = lambda *a: subprocess.check_output(a, stderr=subprocess.DEVNULL).decode("utf-8").strip()
        sha = run("git", "rev-parse", "HEAD")
        branch = run("git", "rev-parse", "--abbrev-ref", "HEAD")
        dirty = subprocess.call(["git", "diff", "--quiet"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) != 0
        return {"sha": sha, "branch": branch, "dirty": dirty}
    except Exception:
        return {"sha": None, "branch": None, "dirty": None}


def _write_json(path: str, payload: dict) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(payload, f, indent=2, sort_keys=True)


def _shell_summary(e_pos, e_neg):
    return {"e_pos": e_pos, "e_neg": e_neg, "e_net": float(e_pos + e_neg), **b200.summary_metrics(e_pos, e_neg)}


def reproduce_main(argv=None) -> int:
    p = argparse.ArgumentParser(description="Reproduce Rodal (2025) irrotational energetics on the B200 engine.")
    p.add_argument("--mode", choices=["axisym", "3d"], default="axisym")
    p.add_argument("--backend", choices=["b200"], default="b200")
    p.add_argument("--dtype", choices=["float64", "float32"], default="float64")
    p.add_argument("--rho", type=float, default=5.0)
    p.add_argument("--sigma", type=float, default=4.0)
    p.add_argument("--v", type=float, default=1.0)
    p.add_argument("--extent-mult", type=float, default=12.0)
    p.add_argument("--extent-pad", type=float, default=1.1)
    p.add_argument("--nx", type=int, default=2400, help="axisym: x grid points")
    p.add_argument("--ny", type=int, default=1200, help="axisym: y grid points")
    p.add_argument("--n", type=int, default=100, help="3d: points per axis")
    p.add_argument("--tail-r1-mult", type=float, default=8.0)
    p.add_argument("--tail-r2-mult", type=float, default=12.0)
    p.add_argument("--out", type=str, default=None, help="Write JSON summary to this path")
    args = p.parse_args(argv)

    extent = args.extent_mult * args.rho * args.extent_pad
    r1 = args.tail_r1_mult * args.rho
    r2 = args.tail_r2_mult * args.rho
    if r2 <= r1:
        raise ValueError("tail-r2-mult must be > tail-r1-mult")

    started = time.perf_counter()
    if args.mode == "axisym":
        run = b200.compute_energy_axisymmetric(rho=args.rho, sigma=args.sigma, v=args.v, extent=extent, nx=args.nx,
                                               ny=args.ny, backend=args.backend, dtype=args.dtype, shells=[r1, r2])
    else:
        run = b200.compute_energy_3d(rho=args.rho, sigma=args.sigma, v=args.v, extent=extent, n=args.n,
                                     backend=args.backend, dtype=args.dtype, shells=[r1, r2])
    e_at_r = run.pop("_e_at_r")
    e_pos_r1, e_neg_r1 = e_at_r(r1)
    e_pos_r2, e_neg_r2 = e_at_r(r2)
    e_pos_inf = b200.two_point_tail_extrapolation(e_pos_r1, e_pos_r2, r1, r2)
    e_neg_inf = b200.two_point_tail_extrapolation(e_neg_r1, e_neg_r2, r1, r2)
    run["counts"]["shell_points"] = {str(k): v for k, v in run["counts"]["shell_points"].items()}
    tail = {"r1": r1, "r2": r2, "at_r1": _shell_summary(e_pos_r1, e_neg_r1), "at_r2": _shell_summary(e_pos_r2, e_neg_r2),
            "at_inf": _shell_summary(e_pos_inf, e_neg_inf)}
    payload = {
        "meta": {"timestamp_utc": datetime.now(timezone.utc).isoformat(), "git": _git_info(), "python": sys.version},
        "params": {"rho": args.rho, "sigma": args.sigma, "v": args.v},
        "run": run,
        "tail_2pt": tail,
        "wall_clock_s": float(time.perf_counter() - started),
    }
    print("-" * 60)
    print(f"Mode: {args.mode}")
    print(f"Params: rho={args.rho}, sigma={args.sigma}, v={args.v}")
    if args.mode == "axisym":
        print(f"Grid: nx={run['grid']['nx']}, ny={run['grid']['ny']}, extent={run['grid']['extent']}")
    else:
        print(f"Grid: n={run['grid']['n']}, extent={run['grid']['extent']}")
    r2_ratio = tail["at_r2"]["ratio_e_pos_over_abs_e_neg"]
    inf_abs, inf_net = tail["at_inf"]["e_abs"], tail["at_inf"]["e_net"]
    # the reference crashes here when e_neg == 0; print n/a instead
    print(f"At R2={r2:.3f}: ratio E+/|E-| = " + (f"{r2_ratio:.6g}" if r2_ratio is not None else "n/a"))
    print("Tail net/abs (∞): " + (f"{abs(inf_net) / inf_abs:.6%}" if inf_abs else "n/a"))
    print("-" * 60)
    if args.out:
        _write_json(args.out, payload)
        print(f"Wrote JSON: {args.out}")
    return 0


def sweep_main(argv=None) -> int:
    import numpy as np
    p = argparse.ArgumentParser(description="Superluminal velocity sweep on the B200 engine")
    p.add_argument("--mode", choices=["axisym", "3d"], default="axisym")
    p.add_argument("--backend", choices=["b200"], default="b200")
    p.add_argument("--dtype", choices=["float64", "float32"], default="float64")
    p.add_argument("--rho", type=float, default=5.0)
    p.add_argument("--sigma", type=float, default=4.0)
    p.add_argument("--v-min", type=float, default=1.0)
    p.add_argument("--v-max", type=float, default=3.0)
    p.add_argument("--v-steps", type=int, default=20)
    p.add_argument("--extent-mult", type=float, default=12.0)
    p.add_argument("--extent-pad", type=float, default=1.1)
    p.add_argument("--nx", type=int, default=1200)
    p.add_argument("--ny", type=int, default=600)
    p.add_argument("--n", type=int, default=80)
    p.add_argument("--tail-r1-mult", type=float, default=8.0)
    p.add_argument("--tail-r2-mult", type=float, default=12.0)
    p.add_argument("--out", type=str, required=True)
    args = p.parse_args(argv)

    v_values = np.linspace(args.v_min, args.v_max, args.v_steps).tolist()
    extent = args.extent_mult * args.rho * args.extent_pad
    started = time.perf_counter()
    sweep_data = sweeps.sweep_velocity(mode=args.mode, rho=args.rho, sigma=args.sigma, v_values=v_values, extent=extent,
                                       tail_r1_mult=args.tail_r1_mult, tail_r2_mult=args.tail_r2_mult,
                                       backend=args.backend, dtype=args.dtype, nx=args.nx, ny=args.ny, n=args.n)
    total_time = time.perf_counter() - started
    payload = {
        "meta": {"timestamp_utc": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "git": _git_info(), "python": sys.version},
        "sweep": sweep_data,
        "total_time_s": total_time,
    }
    if int(os.environ.get("RANK", "0")) == 0:
        _write_json(args.out, payload)
        print(f"Sweep completed in {total_time:.1f}s; results written to: {args.out}")
    return 0


def _array_summary(a):
    """Arrays are summarised as in the reference's io._to_jsonable (io.py:15-21)."""
    import numpy as np
    return {"shape": list(a.shape), "min": float(np.min(a)), "max": float(np.max(a)), "mean": float(np.mean(a))}


def slice_main(argv=None) -> int:
    from .einstein import compute_einstein_eigenvalues
    from .slice2d import compute_slice_z0, integrate_signed
    p = argparse.ArgumentParser(description="z=0 slice diagnostic on the B200 engine (summary JSON of the reference's plot command)")
    p.add_argument("--rho", type=float, default=10.0)
    p.add_argument("--sigma", type=float, default=3.0)
    p.add_argument("--v", type=float, default=1.5)
    p.add_argument("--extent", type=float, default=20.0)
    p.add_argument("--n", type=int, default=301)
    p.add_argument("--einstein", action="store_true", help="Compute Einstein tensor eigenvalues")
    p.add_argument("--tail-correction", action="store_true", help="Compute tail extrapolation for far-field decay")
    p.add_argument("--json-out", type=str, default="results/summary.json")
    args = p.parse_args(argv)
    res = compute_slice_z0(rho=args.rho, sigma=args.sigma, v=args.v, extent=args.extent, n=args.n,
                           tail_correction=args.tail_correction)
    epos, eneg, enet = integrate_signed(res.rho_adm, dx=res.dx, dy=res.dy)
    data = {
        "params": {"rho": args.rho, "sigma": args.sigma, "v": args.v, "extent": args.extent, "n": args.n,
                   "use_einstein": args.einstein, "use_tail_correction": args.tail_correction},
        "integrals_2d": {"E_pos": epos, "E_neg": eneg, "E_net": enet},
        "rho_adm": _array_summary(res.rho_adm),
    }
    t = res.tail_correction
    if t is not None:                                           # viz.py:83-97
        data["tail_correction"] = {
            "exponent": t.exponent, "amplitude": t.amplitude, "fit_r_min": t.fit_r_min, "fit_r_max": t.fit_r_max,
            "tail_integral_pos": t.tail_integral_pos, "tail_integral_neg": t.tail_integral_neg,
            "tail_uncertainty": t.tail_uncertainty, "fit_residual_rms": t.fit_residual_rms,
            "E_pos_corrected": epos + t.tail_integral_pos, "E_neg_corrected": eneg + t.tail_integral_neg,
            "E_net_corrected": (epos + t.tail_integral_pos) - (eneg + t.tail_integral_neg),
        }
    if args.einstein:                                           # viz.py:36-37, 98-103
        er = compute_einstein_eigenvalues(res.beta_x, res.beta_y, dx=res.dx, dy=res.dy)
        data["einstein"] = {"eig_max": _array_summary(er.eig_max), "ricci_scalar": _array_summary(er.ricci_scalar),
                            "type_i_fraction": float(er.is_type_i.sum()) / er.is_type_i.size}
    _write_json(args.json_out, data)
    print(f"Wrote JSON: {args.json_out}")
    return 0


def optimize_main(argv=None) -> int:
    """``python -m irrotational_warp optimize`` (/root/reference/src/irrotational_warp/cli.py:56-77, 179-276): same flags
    and result JSON; ``--objective volume3d`` swaps the 2D-slice objective for |E-| of the 3D exact-potential volume
    (BASELINE config 5: ``--objective volume3d --n 256 --method bayes --n-calls 50``)."""
    from . import optimize as opt
    p = argparse.ArgumentParser(description="Find parameters minimizing negative energy magnitude |E-| on the B200 engine")
    p.add_argument("--rho", type=float, default=10.0, help="Bubble radius (geometric units)")
    p.add_argument("--sigma-min", type=float, default=1.0)
    p.add_argument("--sigma-max", type=float, default=10.0)
    p.add_argument("--sigma-steps", type=int, default=5, help="Grid search resolution (sigma)")
    p.add_argument("--v-min", type=float, default=0.5)
    p.add_argument("--v-max", type=float, default=2.0)
    p.add_argument("--v-steps", type=int, default=5, help="Grid search resolution (v)")
    p.add_argument("--extent", type=float, default=20.0, help="Half-width of the domain")
    p.add_argument("--n", type=int, default=71, help="Grid resolution per axis")
    p.add_argument("--method", type=str, default="hybrid", choices=["grid", "hybrid", "bayes"],
                   help="grid (exhaustive), hybrid (grid+Nelder-Mead), bayes (Bayesian/GP)")
    p.add_argument("--refine", action="store_true", help="[hybrid only] Apply Nelder-Mead refinement after grid search")
    p.add_argument("--n-calls", type=int, default=50, help="[bayes only] Total number of evaluations")
    p.add_argument("--n-initial", type=int, default=10, help="[bayes only] Random evaluations before GP fitting")
    p.add_argument("--random-state", type=int, default=None, help="[bayes only] Random seed for reproducibility")
    p.add_argument("--objective", choices=["slice2d", "volume3d"], default="slice2d",
                   help="slice2d: the reference's z=0 slice objective; volume3d: |E-| of compute_energy_3d")
    p.add_argument("--out", type=str, default="results/optimization.json")
    args = p.parse_args(argv)

    print(f"Running optimization over sigma in [{args.sigma_min}, {args.sigma_max}], v in [{args.v_min}, {args.v_max}]"
          f" ({args.objective} objective)")
    common = dict(rho=args.rho, sigma_range=(args.sigma_min, args.sigma_max), v_range=(args.v_min, args.v_max),
                  extent=args.extent, n=args.n, objective=args.objective)
    if args.method == "bayes":
        print(f"  Bayesian optimization (GP): {args.n_calls} total calls ({args.n_initial} initial random)")
        result = opt.optimize_bayesian(n_calls=args.n_calls, n_initial_points=args.n_initial,
                                       random_state=args.random_state, **common)
    else:
        refine = args.refine and args.method == "hybrid"
        print(f"  Grid search: {args.sigma_steps} x {args.v_steps} = {args.sigma_steps * args.v_steps} evaluations"
              + ("; then Nelder-Mead local refinement" if refine else ""))
        result = opt.optimize_hybrid(sigma_steps=args.sigma_steps, v_steps=args.v_steps, refine=refine, **common)
    params = {"rho": args.rho, "extent": args.extent, "n": args.n, "sigma_range": [args.sigma_min, args.sigma_max],
              "v_range": [args.v_min, args.v_max], "method": args.method, "objective": args.objective}
    if args.method == "bayes":
        params.update(n_calls=args.n_calls, n_initial=args.n_initial)
        if args.random_state is not None:
            params["random_state"] = args.random_state
    else:
        params.update(sigma_steps=args.sigma_steps, v_steps=args.v_steps)
        if args.method == "hybrid":
            params["refine"] = args.refine
    payload = {"git": _git_info(), "params": params,
               "optimization": {"best_params": {k: float(v) for k, v in result.best_params.items()},
                                "best_value": float(result.best_value),
                                "initial_params": {k: float(v) for k, v in result.initial_params.items()},
                                "initial_value": float(result.initial_value), "n_evaluations": int(result.n_evaluations),
                                "success": bool(result.success), "method": result.method, "message": str(result.message)}}
    _write_json(args.out, payload)
    print(f"Optimization complete: best |E-| = {result.best_value:.6e} at sigma = {result.best_params['sigma']:.4f}, "
          f"v = {result.best_params['v']:.4f} ({result.n_evaluations} evaluations, {result.method}); saved to {args.out}")
    return 0


def convergence_main(argv=None) -> int:
    """``scripts/convergence_study_3d.py`` (/root/reference/scripts/convergence_study_3d.py:21-183) on the engine: one
    ``reproduce`` run per resolution, each leaving ``convergence_runs/repro_n{n}.json`` beside ``--out`` as the reference's
    subprocess loop does (here in-process: the engine handle is created once), then the same table and summary JSON."""
    p = argparse.ArgumentParser(description="3D convergence study on the B200 engine")
    p.add_argument("--mode", default="3d", choices=["3d"])
    p.add_argument("--backend", default="b200", choices=["b200"])
    p.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    p.add_argument("--n-values", type=str, default="40,60,80,100", help="Comma-separated list of grid resolutions")
    p.add_argument("--rho", type=float, default=5.0)
    p.add_argument("--sigma", type=float, default=4.0)
    p.add_argument("--v", type=float, default=1.0)
    p.add_argument("--out", type=str, required=True)
    args = p.parse_args(argv)
    n_values = [int(t.strip()) for t in args.n_values.split(",")]
    out_dir = os.path.join(os.path.dirname(args.out) or ".", "convergence_runs")
    os.makedirs(out_dir, exist_ok=True)
    started = time.perf_counter()
    results = []
    for n in n_values:
        out_path = os.path.join(out_dir, f"repro_n{n}.json")
        t0 = time.perf_counter()
        reproduce_main(["--mode", args.mode, "--backend", args.backend, "--dtype", args.dtype, "--n", str(n),
                        "--rho", str(args.rho), "--sigma", str(args.sigma), "--v", str(args.v), "--out", out_path])
        elapsed = time.perf_counter() - t0
        with open(out_path) as f:
            data = json.load(f)
        tail = data["tail_2pt"]["at_inf"]
        results.append({"n": n, "n_points": n ** 3, "e_pos_inf": tail["e_pos"], "e_neg_inf": tail["e_neg"],
                        "e_net_inf": tail["e_net"], "e_abs_inf": tail["e_abs"], "neg_fraction": tail["neg_fraction"],
                        "net_over_abs_pct": tail["net_over_abs"] * 100,
                        "ratio_r2": data["tail_2pt"]["at_r2"]["ratio_e_pos_over_abs_e_neg"],
                        "wall_clock_s": elapsed, "kernel_ms": data["run"]["engine"]["kernel_ms"], "file": out_path})
    total_time = time.perf_counter() - started
    print("=" * 72)
    print(f"{'n':>6} {'Points':>13} {'|E_net/E_abs|%':>15} {'Ratio(R2)':>12} {'Time(s)':>10} {'kernel ms':>10}")
    print("-" * 72)
    for r in results:
        print(f"{r['n']:>6} {r['n_points']:>13,} {abs(r['net_over_abs_pct']):>15.6f} {r['ratio_r2']:>12.5f} "
              f"{r['wall_clock_s']:>10.3f} {r['kernel_ms']:>10.3f}")
    print("=" * 72)
    payload = {"meta": {"timestamp_utc": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "python": sys.version},
               "params": {"mode": args.mode, "backend": args.backend, "dtype": args.dtype, "rho": args.rho,
                          "sigma": args.sigma, "v": args.v},
               "results": results, "total_time_s": total_time}
    _write_json(args.out, payload)
    print(f"Results written to: {args.out}  (total {total_time:.1f}s)")
    return 0


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if argv and argv[0] == "optimize":
        return optimize_main(argv[1:])
    if argv and argv[0] == "convergence":
        return convergence_main(argv[1:])
    if argv and argv[0] == "reproduce":
        return reproduce_main(argv[1:])
    if argv and argv[0] == "sweep":
        return sweep_main(argv[1:])
    if argv and argv[0] == "slice":
        return slice_main(argv[1:])
    print("usage: python -m irrotational_warp_lab_b200.cli {reproduce|sweep|slice|optimize|convergence} [options]", file=sys.stderr)
    return 2


if __name__ == "__main__":
    raise SystemExit(main())

"""patches/reference_b200.patch applied to a copy of the reference: `--backend b200` in the five script front-ends and the
package CLI, and the delegation of the three evaluators to this package.

Needs /root/reference (build container); skipped on the GPU box, where the reference does not exist.  The engine itself
is stubbed: this test is about the seam, the GPU suite is about the numbers."""
import json
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = os.environ.get("IWB_REFERENCE", "/root/reference")
PATCH = os.path.join(ROOT, "patches", "reference_b200.patch")
SCRIPTS = ["reproduce_rodal_exact.py", "sweep_superluminal.py", "convergence_study_3d.py", "profile_3d.py", "compare_sources.py"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="the reference tree is not present")


@pytest.fixture(scope="module")
def patched(tmp_path_factory):
    dst = tmp_path_factory.mktemp("reference_patched")
    shutil.copytree(os.path.join(REF, "scripts"), dst / "scripts")
    shutil.copytree(os.path.join(REF, "src"), dst / "src")
    res = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", PATCH], cwd=dst, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return dst


def _env(patched):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(patched / "src"), str(patched / "scripts"), ROOT])
    return env


def _run(patched, code):
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=patched, env=_env(patched), capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


@pytest.mark.parametrize("script", SCRIPTS)
def test_script_front_ends_offer_b200(patched, script):
    res = subprocess.run([sys.executable, os.path.join("scripts", script), "--help"], cwd=patched, env=_env(patched),
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "b200" in res.stdout and "numpy" in res.stdout


def test_unknown_backend_still_raises_value_error(patched):
    out = _run(patched, """
        import reproduce_rodal_exact as r
        try:
            r._resolve_backend("tpu")
        except ValueError as e:
            print("ValueError", e)
    """)
    assert "ValueError" in out and "numpy|cupy|b200" in out


def test_evaluators_delegate_to_the_engine_package(patched):
    """compute_energy_3d / compute_energy_axisymmetric(backend="b200") hand their arguments to
    irrotational_warp_lab_b200.backend unchanged and return its dictionary; numpy stays numpy."""
    out = _run(patched, """
        import json
        import reproduce_rodal_exact as r
        from irrotational_warp_lab_b200 import backend as b200
        calls = []
        b200.compute_energy_3d = lambda **kw: calls.append(("3d", kw)) or {"mode": "3d", "stub": True}
        b200.compute_energy_axisymmetric = lambda **kw: calls.append(("axisym", kw)) or {"mode": "axisym", "stub": True}
        a = r.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40, backend="b200", dtype="float64")
        b = r.compute_energy_axisymmetric(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=60, ny=30, backend="B200", dtype="float32")
        c = r.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=12, backend="numpy", dtype="float64")
        print(json.dumps({"a": a, "b": b, "calls": calls, "c_backend": c["backend"]}))
    """)
    d = json.loads(out)
    assert d["a"] == {"mode": "3d", "stub": True} and d["b"] == {"mode": "axisym", "stub": True}
    assert d["calls"][0] == ["3d", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40, backend="b200", dtype="float64")]
    assert d["calls"][1] == ["axisym", dict(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=60, ny=30, backend="b200", dtype="float32")]
    assert d["c_backend"] == "numpy"


def test_reference_main_runs_on_the_engine_dictionary(patched, tmp_path):
    """The reference's own main() -- tail extrapolation, prints, JSON payload -- consumes the dictionary shape of
    backend.compute_energy_3d (timing_s.phi is None, extra `counts` / `engine` blocks, callable _e_at_r).  The engine is
    replaced by the NumPy oracle so that the numbers are the reference's."""
    out_json = tmp_path / "run.json"
    _run(patched, f"""
        import reproduce_rodal_exact as r
        from irrotational_warp_lab_b200 import backend as b200
        from oracle import oracle_np

        class FakeEngine:
            name = "stub"
            def energy3d(self, *, x, y, z, dx, dy, dz, rho, sigma, v, dtype, shells, eps, rho_out, mode):
                o = oracle_np.energy_3d(rho=rho, sigma=sigma, v=v, extent=float(-x[0]), n=len(x), dtype=dtype, shells=shells)
                return dict(e_pos=o["e_pos"], e_neg=o["e_neg"], cnt_pos=o["cnt_pos"], cnt_neg=o["cnt_neg"], cnt_zero=o["cnt_zero"],
                            cnt_nan=0, shell_pos=[s["e_pos"] for s in o["shells"]], shell_neg=[s["e_neg"] for s in o["shells"]],
                            shell_cnt=[s["cnt"] for s in o["shells"]], kernel_ms=0.0, total_ms=0.0)
        b200.resolve_backend = lambda name: (FakeEngine(), "b200")
        assert r.main(["--mode", "3d", "--backend", "b200", "--n", "40", "--out", r"{out_json}"]) == 0
    """)
    payload = json.load(open(out_json))
    run = payload["run"]
    assert run["backend"] == "b200" and run["mode"] == "3d" and run["timing_s"]["phi"] is None
    # the reference's n = 40 numbers
    assert run["energies"]["e_pos"] == 0.6835753113559453 and run["energies"]["e_neg"] == -0.6324831088550202
    assert abs(payload["tail_2pt"]["at_r2"]["ratio_e_pos_over_abs_e_neg"] - 1.113406636389579) < 1e-12
    assert "_e_at_r" not in run and run["counts"]["neg"] == 46400


def test_package_cli_backend_flag_reaches_the_slice_evaluator(patched):
    """`python -m irrotational_warp --backend b200 ...`: the parser offers the flag and adm.compute_slice_z0 -- and through
    it sweep.py and optimize.py, unchanged -- evaluates on irrotational_warp_lab_b200.slice2d.  matplotlib is absent in
    the build container, so the two modules cli.py imports for plotting are stubbed."""
    out = _run(patched, """
        import sys, types, json
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
            m = types.ModuleType(name); m.use = lambda *a, **k: None; sys.modules[name] = m
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        from irrotational_warp import adm, cli, sweep, optimize
        from irrotational_warp_lab_b200 import slice2d
        import numpy as np
        help_text = cli.build_parser().format_help()
        calls = []
        def fake(**kw):
            calls.append(kw)
            z = np.zeros((kw["n"], kw["n"]))
            return slice2d.SliceResult(x=np.zeros(kw["n"]), y=np.zeros(kw["n"]), rho_adm=z - 1.0, phi=z, beta_x=z, beta_y=z, dx=0.5, dy=0.5)
        slice2d.compute_slice_z0 = fake
        args = cli.build_parser().parse_args(["--backend", "b200", "sweep", "--n", "9"])
        adm.set_backend(args.backend)
        pts = sweep.sweep_sigma_z0(rho=10.0, v=1.5, sigma_values=[2.0, 3.0], extent=20.0, n=9)
        val = optimize.objective_neg_energy(np.array([3.0, 1.5]), ["sigma", "v"], 10.0, 20.0, 9)
        adm.set_backend("numpy")
        ref = adm.compute_slice_z0(rho=10.0, sigma=3.0, v=1.5, extent=20.0, n=9)
        try:
            adm.set_backend("cupy")
            bad = False
        except ValueError:
            bad = True
        print(json.dumps({"help_has_flag": "--backend" in help_text and "b200" in help_text, "ncalls": len(calls),
                          "e_neg": [p.e_neg for p in pts], "val": val, "numpy_type": type(ref).__module__, "bad": bad,
                          "sigmas": [c["sigma"] for c in calls]}))
    """)
    d = json.loads(out)
    assert d["help_has_flag"] and d["ncalls"] == 3 and d["sigmas"] == [2.0, 3.0, 3.0]
    assert d["e_neg"] == [81 * 0.25, 81 * 0.25] and d["val"] == 81 * 0.25      # 81 points of rho = -1, dA = 0.25
    assert d["numpy_type"] == "irrotational_warp.adm" and d["bad"]

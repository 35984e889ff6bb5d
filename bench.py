#!/usr/bin/env python3
"""Headline benchmark: 3D energy-condition evaluation, fp64, n = 1024^3 (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1024] [--impl ours|reference]

One *step* = one full evaluation of the hot path on the synthetic (analytic) grid rho=5 sigma=4 v=1 extent=66, shells
r<=40 / r<=60.  One GPU: fused kernel -> fixed-order tile reduction -> fixed-order row reduction.  N > 1 (torchrun, one
rank per GPU): every rank marches an interleaved shard of the (j, k) tiles along the whole of axis 0, pushes its chain
sums into every rank's gathered buffer as NVLink peer stores, and combines them with the single-GPU tree (bit-identical
energies on every rank; the NCCL all-gather is the fall-back when the ranks cannot map each other's memory).  The
volume is fixed, so "scaling": "strong".

Printed JSON (one line, rank 0):
  value   : Gpoints/s, whole job, coordinates resident in HBM -- per-step CUDA-event pairs on the launching stream,
            summed over K steps, max over ranks.  For every N the result row of each timed step stays on the device;
            all K rows are read after the closing barrier and must equal the warm-up's result.
  e2e     : same metric through the public API (backend.compute_energy_3d / distributed.compute_energy_3d_sharded):
            host numpy coordinates staged through pinned memory and the 240-byte result read back EVERY step, wall clock
  parity  : the run's result against the frozen full-volume answer (tests/golden/golden_3d_large.json): energies
            <= 1e-10 relative, sign counts and shell point counts exact -- asserted, not just reported
  multi_gpu_parity (N > 1): slab-mode, NCCL-exchange, unseen-radius and rank-dealt sweep results against rank 0's
            single-GPU answers, bit for bit
  roofline: FP64-pipe bound (no dense contraction, ~0 HBM bytes/point); see DESIGN.md section 6
  cpu_baseline: the NumPy restatement of the reference (oracle/oracle_np.py) on every host core, bounded sample
`--impl reference` times that CPU path alone with the same --steps / --warmup, on every host core (worker processes over
slabs of the sample volume; the reference itself is single-threaded NumPy and needs ~240 GB at n = 1024^3).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RHO, SIGMA, V, EXTENT = 5.0, 4.0, 1.0, 66.0
SHELLS = [40.0, 60.0]
METRIC = "3D Gpoints/s (fp64, n=1024^3)"
# FP64 issue slots per grid point.  CONTRACT = SURVEY.md 8(d) / BASELINE.md 4 (IEEE divisions at 9 slots each, exp/log1p
# on every point).  EXACT_ALG = the bit-exact algorithm this engine runs with every grid point evaluated once (no halo
# recompute): phi 32 (sqrt 8, division with a seeded reciprocal 4, the other 5, remaining arithmetic 15), nine
# derivatives x 4 = 36, rho * dV 19, sums 4 (round 1 counted 95 with two 6-slot divisions and a
# two-addition row test).
SLOTS_CONTRACT = 271.0
SLOTS_EXACT_ALG = 91.0
# per-launch figures of k_energy3d at n = 1024^3 on one GPU from the committed `ncu --set full` capture of this build
# (profiles/r2_k_energy3d_n1024_summary.txt): executed FP64 thread instructions per point, DRAM bytes read + written
NCU_CAPTURE = "profiles/r2_k_energy3d_n1024_summary.txt"
NCU_FP64_EXECUTED_PER_POINT = 102.6
NCU_DRAM_BYTES_PER_LAUNCH = 110848 + 0
CPU_SAMPLE_N = 256
FP64_LANES_PER_SM = 64


def workload_config(n):
    """The `config` block: identical in both arms (the driver compares them)."""
    return {
        "workload": f"3D n={n}^3 fp64 energy-condition evaluation (BASELINE config 4): rho=5 sigma=4 v=1 extent=66, "
                    f"shells r<=40 and r<=60, exact (IEEE-replay) mode",
        "n": n, "rho": RHO, "sigma": SIGMA, "v": V, "extent": EXTENT, "shells": SHELLS, "dtype": "float64", "mode": "exact",
        "l2": "GPU arm: a 256 MiB buffer is written between timed steps (the inputs are 32 KB of coordinates); "
              "CPU arm: not applicable",
        "timing": "GPU arm: per-step CUDA-event pairs on the launching stream, summed, max over ranks; CPU arm: wall clock",
    }


def load_golden_case(name):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "golden_3d_large.json")) as f:
            return json.load(f)["cases"].get(name)
    except OSError:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, smax, reasons = [], [], set()
        # the sampling period is 100 ms and an 8-GPU timed region is ~30 ms: take the samples that bracket it as well
        # (the GPU runs the same steps in the warm-up just before and in the kernel-only / end-to-end legs just after)
        for t, line in self.rows:
            if t < t0 - 0.12 or t > t1 + 0.12:
                continue
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the NumPy restatement of the reference on every host core
# ---------------------------------------------------------------------------------------------------------------------
def _reference_slab(bounds):
    """Worker: the NumPy port on planes [a, b) of the sample volume (its 2-plane phi halo recomputed)."""
    from oracle import oracle_np
    a, b = bounds
    r = oracle_np.energy_3d(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=CPU_SAMPLE_N, shells=SHELLS, i_begin=a, i_end=b)
    return r["e_pos"], r["e_neg"]


def _reference_timed(steps, warmup):
    """`warmup` untimed + `steps` timed evaluations of the sample volume (n = CPU_SAMPLE_N), cut into slabs of axis 0 that
    worker processes evaluate independently -- the reference's own convergence_study_3d.py fans out over subprocesses
    in the same way; the reference itself is single-threaded NumPy.  Returns seconds per step and the worker count."""
    import multiprocessing as mp
    n = CPU_SAMPLE_N
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = max(1, min(avail, 32, n // 8))
    step = -(-n // cores)
    slabs = [(a, min(n, a + step)) for a in range(0, n, step)]
    ctx = mp.get_context("spawn")      # the main arm has CUDA and helper threads alive: never fork it
    with ctx.Pool(processes=cores) as pool:
        for _ in range(max(warmup, 1)):
            pool.map(_reference_slab, slabs)
        t0 = time.perf_counter()
        for _ in range(steps):
            parts = pool.map(_reference_slab, slabs)
        dt = (time.perf_counter() - t0) / steps
    e_pos, e_neg = sum(p[0] for p in parts), sum(p[1] for p in parts)
    # n = 256: the reference gives 1.0274931410661474 / -0.9752262663519845 (tests/golden/golden_3d.json)
    assert abs(e_pos - 1.0274931410661474) < 1e-12 and abs(e_neg + 0.9752262663519845) < 1e-12, (e_pos, e_neg)
    return dt, cores, avail, len(slabs)


def _reference_record(dt, cores, avail, nslabs, n_full):
    n = CPU_SAMPLE_N
    return {"value": n ** 3 / dt / 1e9, "unit": "Gpoints/s", "cores": cores, "kind": "port",
            "sample_n": n,
            "sample": f"each step = one n={n}^3 evaluation, a bounded sample of the n={n_full}^3 workload (the reference needs "
                      f"~224 B/point, i.e. ~240 GB at 1024^3; its points/s does not depend on n), as {nslabs} slabs of axis 0 "
                      f"on {cores} worker processes; NumPy port oracle/oracle_np.py, bit-equal to the reference on the golden "
                      f"grids (the reference itself is single-threaded; /root/reference does not exist on the GPU box)",
            "host_cores_available": avail}


def cpu_baseline(n_full=1024, reps=3):
    """cpu_baseline leg of the main arm: the same measurement as `--impl reference`, a few steps."""
    return _reference_record(*_reference_timed(reps, 1), n_full)


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host, every core, --steps / --warmup honoured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    dt, cores, avail, nslabs = _reference_timed(steps, warmup)
    rec = _reference_record(dt, cores, avail, nslabs, args.n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": "Gpoints/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n),
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def check_parity(res, n):
    """The run's result against the frozen answer for this grid (None when no golden exists for n)."""
    rec = load_golden_case(f"n{n}")
    if rec is None:
        return None
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)      # noqa: E731
    out = {
        "against": f"tests/golden/golden_3d_large.json:n{n} ({rec['source']}"
                   + ("; the reference cannot run this size, the generating script validated the C oracle against the "
                      "reference's n=400 and n=512 runs)" if rec["source"] != "reference" else ")"),
        "e_pos_rel_err": rel(res["e_pos"], rec["energies"]["e_pos"]),
        "e_neg_rel_err": rel(res["e_neg"], rec["energies"]["e_neg"]),
        "counts_exact": (res["cnt_pos"], res["cnt_neg"], res["cnt_zero"]) == (rec["cnt_pos"], rec["cnt_neg"], rec["cnt_zero"]),
        "shell_counts_exact": list(res["shell_cnt"]) == [s["cnt"] for s in rec["shells"]],
        "shell_rel_err": max(max(rel(res["shell_pos"][k], s["e_pos"]), rel(res["shell_neg"][k], s["e_neg"]))
                             for k, s in enumerate(rec["shells"])),
        "tolerance": 1e-10,
    }
    out["ok"] = bool(out["e_pos_rel_err"] <= 1e-10 and out["e_neg_rel_err"] <= 1e-10 and out["shell_rel_err"] <= 1e-10
                     and out["counts_exact"] and out["shell_counts_exact"])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: 20-point superluminal sweep + 50-point optimiser batch, 3D n = 256, spread over the GPUs
# ---------------------------------------------------------------------------------------------------------------------
def run_sweep_opt(args):
    """`--config sweep_opt`: one step = the 20 sweep evaluations of sweeps.sweep_velocity (v = 1..3) plus the 50 objective
    evaluations of optimize.evaluate_objective_3d_batch ((sigma, v) drawn with seed 42 from [2, 8] x [0.8, 2.0], the box and
    budget of the reference README's Bayesian run), all 3D n = 256^3, rho = 5.  N > 1: both calls deal their points
    round-robin to the ranks.  Timed through the public API with host buffers: wall clock between barriers, max over ranks."""
    import torch
    import torch.distributed as dist

    from irrotational_warp_lab_b200 import optimize, sweeps
    from irrotational_warp_lab_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    get_engine(local_rank)
    n = 256
    v_values = np.linspace(1.0, 3.0, 20).tolist()
    gold = load_golden_case("objective_n256")
    rng = np.random.default_rng(42)
    pts = [(p["sigma"], p["v"]) for p in gold["points"]] if gold else []
    pts += [(float(s), float(v)) for s, v in zip(rng.uniform(2.0, 8.0, 50 - len(pts)), rng.uniform(0.8, 2.0, 50 - len(pts)))]
    dkw = dict(distributed=True) if world > 1 else {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        sw = sweeps.sweep_velocity(mode="3d", rho=RHO, sigma=SIGMA, v_values=v_values, extent=EXTENT, tail_r1_mult=8.0,
                                   tail_r2_mult=12.0, n=n, **dkw)
        vals = optimize.evaluate_objective_3d_batch(pts, RHO, EXTENT, n, **dkw)
        return sw, vals

    warmup, steps = max(args.warmup, 3), max(args.steps, 1)
    for _ in range(warmup):
        sw, vals = step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        sw, vals = step()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    nev = len(v_values) + len(pts)
    # parity: the frozen reference objective values, and E proportional to v^2 along the sweep (to 1e-15)
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)      # noqa: E731
    obj_err = max((rel(vals[k], p["abs_e_neg"]) for k, p in enumerate(gold["points"])), default=None) if gold else None
    e1 = sw["sweep_results"][0]["e_pos_inf"]
    v2_err = max(rel(r["e_pos_inf"], e1 * r["v"] ** 2) for r in sw["sweep_results"])
    assert (obj_err is None or obj_err < 1e-10) and v2_err < 1e-12, (obj_err, v2_err)
    if rank == 0:
        from oracle import oracle_np
        t1 = time.perf_counter()
        oracle_np.energy_3d(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS)
        t_cpu = time.perf_counter() - t1
        print(json.dumps({
            "metric": "config-5 evaluations/s (20-point v sweep + 50-point optimiser batch, 3D n=256^3)",
            "value": nev * steps / dt, "unit": "evaluations/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: sweeps.sweep_velocity (v = 1..3, 20 points) + "
                                   "optimize.evaluate_objective_3d_batch (50 (sigma, v) points, seed 42, sigma in [2, 8], v in "
                                   "[0.8, 2.0]), 3D n=256^3, rho=5, exact mode; points dealt round-robin to the ranks",
                       "evaluations_per_step": nev, "n": n},
            "gpoints_per_s": nev * steps * n ** 3 / dt / 1e9,
            "timing": "public API, host buffers in / Python floats out every evaluation; wall clock between barriers, max over ranks",
            "parity": {"objective_rel_err_vs_reference_golden": obj_err, "sweep_v2_scaling_rel_err": v2_err},
            "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "evaluations/s", "cores": 1, "kind": "port",
                             "sample": "one n=256^3 evaluation of the NumPy port (the reference evaluates the points serially, "
                                       "single-threaded)"},
        }))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 2: reproduce_rodal_exact --mode axisym, nx = 2400, ny = 1200, one B200
# ---------------------------------------------------------------------------------------------------------------------
def run_axisym(args):
    """`--config axisym`: one step = one backend.compute_energy_axisymmetric call (host linspace coordinates in, Python
    floats out), wall clock; the kernel's CUDA-event time beside it; parity against the reference's frozen result."""
    from irrotational_warp_lab_b200 import backend
    nx, ny = 2400, 1200
    kw = dict(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, nx=nx, ny=ny)
    warmup, steps = max(args.warmup, 3), max(args.steps, 1)
    for _ in range(warmup):
        run = backend.compute_energy_axisymmetric(**kw)
    kms = []
    t0 = time.perf_counter()
    for _ in range(steps):
        run = backend.compute_energy_axisymmetric(**kw)
        kms.append(run["engine"]["kernel_ms"])
    dt = (time.perf_counter() - t0) / steps
    with open(os.path.join(ROOT, "tests", "golden", "golden_axisym.json")) as f:
        gold = json.load(f)["cases"]["axi_2400x1200"]
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-300)      # noqa: E731
    e_at_r = run["_e_at_r"]
    shell_err = max(max(rel(e_at_r(s["r"])[0], s["e_pos"]), rel(e_at_r(s["r"])[1], s["e_neg"])) for s in gold["shells"])
    parity = {"against": "tests/golden/golden_axisym.json:axi_2400x1200 (reference)",
              "e_pos_rel_err": rel(run["energies"]["e_pos"], gold["energies"]["e_pos"]),
              "e_neg_rel_err": rel(run["energies"]["e_neg"], gold["energies"]["e_neg"]),
              "shell_rel_err": shell_err,
              "counts_exact": (run["counts"]["pos"], run["counts"]["neg"], run["counts"]["zero"]) ==
                              (gold["cnt_pos"], gold["cnt_neg"], gold["cnt_zero"])}
    assert parity["counts_exact"] and max(parity["e_pos_rel_err"], parity["e_neg_rel_err"], shell_err) < 1e-10, parity
    from oracle import oracle_np
    t1 = time.perf_counter()
    oracle_np.energy_axisym(shells=SHELLS, **kw)
    t_cpu = time.perf_counter() - t1
    pts = nx * ny
    print(json.dumps({
        "metric": "axisym Gpoints/s (fp64, nx=2400 ny=1200)", "value": pts / dt / 1e9, "unit": "Gpoints/s", "n_gpus": 1,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "none",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: compute_energy_axisymmetric rho=5 sigma=4 v=1 extent=66 nx=2400 ny=1200, "
                               "shells r<=40 and r<=60, through the public API (host buffers in, Python floats out)"},
        "kernel_ms": statistics.median(kms), "kernel_gpoints_per_s": pts / (statistics.median(kms) * 1e-3) / 1e9,
        "parity": parity,
        "cpu_baseline": {"value": pts / t_cpu / 1e9, "unit": "Gpoints/s", "cores": 1, "kind": "port",
                         "sample": "one evaluation of the NumPy port at the same size (the reference is single-threaded)"},
    }))
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["headline", "sweep_opt", "axisym"], default="headline",
                    help="headline: BASELINE config 4 (the driver's line); sweep_opt: BASELINE config 5; axisym: config 2")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--exchange", choices=["auto", "peer", "nccl"], default="auto",
                    help="N > 1: how the chain sums travel (NVLink peer stores fused behind the kernel, or NCCL all-gather)")
    ap.add_argument("--shard", choices=["tiles", "slab"], default="tiles",
                    help="N > 1: interleaved tile shards (default) or slabs of axis 0 (the north star's wording; 2-plane halo "
                         "recomputed analytically, NCCL all-gather of the table rows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args(argv)
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "sweep_opt":
        return run_sweep_opt(args)
    if args.config == "axisym":
        return run_axisym(args)

    import torch
    import torch.distributed as dist

    from irrotational_warp_lab_b200 import _capi, backend, distributed, sharding, sweeps
    from irrotational_warp_lab_b200._capi import ROW_WIDTH, TILE_CHAINS
    from irrotational_warp_lab_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if _capi.LIB_IS_OVERRIDDEN:
        raise SystemExit("bench.py measures the product library: unset IWB_LIB / IWB_ALLOW_LIB_OVERRIDE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the b200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    n = args.n
    eng = get_engine(local_rank)
    x = np.linspace(-EXTENT, EXTENT, n)
    dx = float(x[1] - x[0])
    # N > 1: interleaved tile shards (every rank marches its tiles along the whole of axis 0); the slab decomposition
    # is the fall-back for a world size that does not divide the 16 tile chains
    tiles = sharding.tile_shard(world, rank) if (world > 1 and args.shard == "tiles") else None
    i_begin, i_end = (0, n) if (world == 1 or tiles is not None) else sharding.slab_bounds(n, world)[rank]
    nfb = sharding.num_flush_blocks(n)
    r0, r1 = sharding.rows_of_slab(i_begin, i_end)
    peers = None
    if tiles is not None and args.exchange != "nccl":
        peers = distributed.peer_group_for(eng, None, nfb)
        if peers is None and args.exchange == "peer":
            raise SystemExit("--exchange peer: the ranks cannot map each other's device memory")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: coordinates resident on the device, CUDA events ----------------------------
    shard_kw = dict(tile_stride=tiles[0], tile_first=tiles[1]) if tiles is not None else {}
    plan = eng.plan3d(x=x, y=x, z=x, dx=dx, dy=dx, dz=dx, rho=RHO, sigma=SIGMA, v=V, dtype="float64", shells=SHELLS,
                      i_begin=i_begin, i_end=i_end, **shard_kw) if i_end > i_begin else None
    if tiles is not None:
        table = torch.zeros((nfb, TILE_CHAINS // world, ROW_WIDTH), dtype=torch.float64, device=dev)     # chain sums
    else:
        table = torch.zeros((nfb, ROW_WIDTH), dtype=torch.float64, device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    # nvidia-smi is started BEFORE the warm-up and given time to initialise NVML: its start-up briefly stalls the driver,
    # which showed up as 0.2-0.3 ms outliers in the 1.4 ms steps of an 8-GPU run when it fell into the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    # Every step leaves its result row on the device (rows_out[s]); nothing is read back inside the timed region, so no
    # host thread sits inside the per-step event pairs -- the same flavour for every N.
    rows_out = torch.zeros((warmup + steps, ROW_WIDTH), dtype=torch.float64, device=dev)
    scratch_tab = torch.empty((nfb, ROW_WIDTH), dtype=torch.float64, device=dev)
    keep_alive = []

    def resident_step(s):
        row = rows_out[s].data_ptr()
        if peers is not None:
            plan.run_exchange_async(peers, row, stream.cuda_stream)          # k_energy3d, push over NVLink, combine
            return
        if plan is not None:
            plan.run_async(table.data_ptr(), stream.cuda_stream)
        if tiles is not None:
            full = sharding.all_gather_chains(table, world)
            keep_alive.append(full)
            eng.reduce_chains_async(full.data_ptr(), world, nfb, scratch_tab.data_ptr(), row, stream.cuda_stream)
        else:
            full = sharding.all_gather_table(table[r0:r1], n, world) if world > 1 else table
            keep_alive.append(full)
            eng.reduce_table_async(full.data_ptr(), nfb, row, stream.cuda_stream)

    for s in range(warmup):
        resident_step(s)
    torch.cuda.synchronize(dev)
    res = eng.read_row(rows_out[warmup - 1].data_ptr(), len(SHELLS), stream.cuda_stream)     # what the timed steps must reproduce
    res["cnt_nan"] = n ** 3 - res["cnt_pos"] - res["cnt_neg"] - res["cnt_zero"]
    gc.collect()
    gc.disable()          # no collector pauses between the launches of the timed steps (one late rank delays every rank)
    barrier()
    t_region0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s in range(steps):
        flush_buf.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        resident_step(warmup + s)
        ev[s][1].record(stream)
    barrier()
    t_region1 = time.perf_counter()
    rows_host = rows_out.cpu().numpy()
    for s in range(warmup + steps):
        got = sharding.reduce_table_host(rows_host[s:s + 1], len(SHELLS))
        assert (got["e_pos"], got["e_neg"], got["cnt_neg"], got["cnt_pos"]) == \
               (res["e_pos"], res["e_neg"], res["cnt_neg"], res["cnt_pos"]), f"step {s} disagrees with the warm-up"
    keep_alive.clear()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(step_ms)
    if os.environ.get("IWB_BENCH_VERBOSE"):
        print(f"rank {rank} step ms: " + " ".join(f"{t:.3f}" for t in step_ms), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / steps
    value = n ** 3 / (ms_per_step * 1e-3) / 1e9
    # own kernels per step: k_energy3d + (peer: k_chain_sums_push, k_exchange_combine | NCCL: k_chain_sums,
    # k_combine_chains, k_reduce_rows | one GPU: k_reduce_tiles, k_reduce_rows)
    per_step = (1 if plan is not None else 0) + (2 if (peers is not None or tiles is None) else 3)
    launches_value = (warmup + steps) * per_step

    # kernel-only duration of the dominant kernel (this rank's share), CUDA events on its launching stream
    kern_ms = []
    if plan is not None:
        for _ in range(min(steps, 5)):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            plan.run_async(table.data_ptr(), stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize(dev)
            kern_ms.append(a.elapsed_time(b))
    kern_ms_avg = sum(kern_ms) / max(len(kern_ms), 1)
    launches_kernel_only = 2 * len(kern_ms)

    # The clock sampler covers the device-timed legs only: its 100 ms nvidia-smi polling stalls the driver for a moment at
    # every tick, which blocking calls (one synchronisation per step) feel directly -- 0.3 ms per 9.4 ms step in round 2's
    # first runs -- while the queued launches of the legs above do not.
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_region0, time.perf_counter())

    # ---------------- e2e: public API, host buffers, wall clock ------------------------------------------
    def api_step():
        if world > 1:
            return distributed.compute_energy_3d_sharded(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS,
                                                         exchange=args.exchange, shard=args.shard)
        return backend.compute_energy_3d(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS)

    for _ in range(warmup):
        run = api_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run = api_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n ** 3 * steps / e2e_s / 1e9
    launches_e2e = (warmup + steps) * (3 if (world == 1 or peers is not None) else 4)
    gc.enable()

    # the two paths agree bit for bit, and with the frozen full-volume answer
    assert run["energies"]["e_pos"] == res["e_pos"] and run["energies"]["e_neg"] == res["e_neg"], "paths disagree"
    assert run["counts"]["neg"] == res["cnt_neg"] and run["counts"]["pos"] == res["cnt_pos"], "paths disagree (counts)"
    parity = check_parity(res, n)
    if parity is not None:
        assert parity["ok"], parity

    # ---------------- N > 1: the other multi-process paths against rank 0's single-GPU answers --------------------
    multi = None
    if world > 1:
        multi = {}
        r_extra = 23.5
        single = backend.compute_energy_3d(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS) if rank == 0 else None
        single_extra = single["_e_at_r"](r_extra) if rank == 0 else None

        def same_as_single(got, got_extra):
            if rank != 0:
                return True
            return bool(got["energies"]["e_pos"] == single["energies"]["e_pos"] and got["energies"]["e_neg"] == single["energies"]["e_neg"]
                        and got["counts"] == single["counts"] and tuple(got_extra) == tuple(single_extra))

        for label, kw in (("tiles_" + ("peer" if peers is not None else "nccl"), dict(shard="tiles", exchange=args.exchange)),
                          ("tiles_nccl", dict(shard="tiles", exchange="nccl")),
                          ("slab_nccl", dict(shard="slab"))):
            got = distributed.compute_energy_3d_sharded(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS, **kw)
            extra = got["_e_at_r"](r_extra)            # a radius nobody asked for: collective relaunch
            multi[label + "_equals_single_gpu"] = same_as_single(got, extra)
        vs = [1.0, 1.5, 2.0, 2.5, 3.0]
        sw = sweeps.sweep_velocity(mode="3d", rho=RHO, sigma=SIGMA, v_values=vs, extent=EXTENT, tail_r1_mult=8.0,
                                   tail_r2_mult=12.0, n=128, distributed=True)          # points dealt round-robin to the ranks
        if rank == 0:
            sw1 = sweeps.sweep_velocity(mode="3d", rho=RHO, sigma=SIGMA, v_values=vs, extent=EXTENT, tail_r1_mult=8.0,
                                        tail_r2_mult=12.0, n=128)
            key = lambda d: [(r["v"], r["e_pos_inf"], r["e_neg_inf"], r["ratio_r2"]) for r in d["sweep_results"]]   # noqa: E731
            multi["rank_dealt_sweep_equals_single_gpu"] = bool(key(sw) == key(sw1))
            multi["unseen_radius"] = r_extra
            assert all(v for k, v in multi.items() if k.endswith("_equals_single_gpu")), multi
        barrier()

    if rank == 0:
        peak = eng.measure_fma_peak(False)
        kernel_points = (i_end - i_begin) * n * n // (world if tiles is not None else 1)     # this rank's share
        pts_per_s_kernel = kernel_points / (kern_ms_avg * 1e-3) if kern_ms_avg else 0.0
        achieved_contract = pts_per_s_kernel * SLOTS_CONTRACT * 2 / 1e12
        achieved_exact = pts_per_s_kernel * SLOTS_EXACT_ALG * 2 / 1e12
        sm_max_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_nominal = eng.sm_count * FP64_LANES_PER_SM * 2 * sm_max_mhz * 1e6 / 1e12
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        # coordinates in; per-(flush block, tile) partial rows out, consumed by the tile reduction and usually never leaving L2
        ntiles = _capi.load().iwb_debug_tiles_j(n, None, 0) * _capi.load().iwb_debug_tiles_k(n, None, 0)
        alg_bytes = 8.0 * (2 * n + n + n) + 8.0 * ROW_WIDTH * (r1 - r0) * ntiles / (world if tiles is not None else 1)
        cfg = workload_config(n)
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "sharding": ("single GPU" if world == 1 else
                         f"{world} interleaved tile shards (every {world}-th 36x60 tile of the (j,k) plane, whole axis 0); "
                         + (f"{nfb}x{TILE_CHAINS}x{ROW_WIDTH} f64 chain sums pushed into every rank's buffer as NVLink peer stores "
                            f"(k_chain_sums_push + k_exchange_combine, no NCCL launch)" if peers is not None else
                            f"NCCL all-gather of {nfb}x{TILE_CHAINS}x{ROW_WIDTH} f64 chain sums") if tiles is not None else
                         f"{world} slab(s) of axis 0, analytic halo recompute, all-gather of a {nfb}x{ROW_WIDTH} f64 table"),
            "e2e": {"value": e2e_value, "unit": "Gpoints/s", "h2d_bytes_per_step": 8 * 4 * n,
                    "d2h_bytes_per_step": 8 * ROW_WIDTH, "ms_per_step": e2e_s / steps * 1e3,
                    "api": "distributed.compute_energy_3d_sharded" if world > 1 else "backend.compute_energy_3d"},
            "gpu_launches": launches_value + launches_kernel_only + launches_e2e,
            "parity": parity,
            "roofline": {
                "bound": "fp64", "kernel": "k_energy3d", "unit": "TFLOP/s",
                "achieved": achieved_exact, "peak": peak["tflops"], "frac": achieved_exact / peak["tflops"],
                "peak_nominal": peak_nominal, "frac_nominal": achieved_exact / peak_nominal,
                "peak_source": "register-resident DFMA microbenchmark run in this process (iwb_measure_fma_peak: one CTA of 16 "
                               "warps per SM, 8 chains per thread; the pipe is 100 % busy, so this is lanes x the SM clock it ran "
                               "at); MEASURED_PEAKS.json has no FP64 vector figure.  peak_nominal = SMs x 64 lanes x 2 x "
                               "clocks.max.sm",
                "peak_sm_mhz": peak["sm_mhz_effective"],
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (n == 1024 and world == 1) else None,
                "traffic_source": NCU_CAPTURE if (n == 1024 and world == 1) else None,
                "slots_per_point": SLOTS_EXACT_ALG,
                "fp64_executed_per_point": NCU_FP64_EXECUTED_PER_POINT if n == 1024 else None,
                "note": "FP64-pipe bound: achieved = kernel points/s x 91 FP64 issue slots/point x 2 -- the slots of the "
                        "bit-exact algorithm with every point evaluated once (DESIGN.md section 6); the kernel executes more "
                        "(fp64_executed_per_point, from the ncu capture named in traffic_source: phi is recomputed in the "
                        "tile halos), which does not count as achieved work",
                "kernel_ms": kern_ms_avg, "kernel_points": kernel_points,
                "contract": {"slots_per_point": SLOTS_CONTRACT, "achieved": achieved_contract,
                             "frac": achieved_contract / peak["tflops"],
                             "note": "SURVEY.md 8(d) figure (9-slot IEEE divisions, exp/log1p at every point)"},
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms_avg * 1e-3) / 1e9
                        if kern_ms_avg else None, "peak_gbs": hbm_peak, "note": "never binding: ~0 B/point"},
            },
            "clocks": clocks,
            "result": {"e_pos": res["e_pos"], "e_neg": res["e_neg"], "cnt_neg": res["cnt_neg"], "cnt_pos": res["cnt_pos"],
                       "cnt_zero": res["cnt_zero"], "shell_cnt": list(res["shell_cnt"])},
        }
        if multi is not None:
            line["multi_gpu_parity"] = multi
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
        print(json.dumps(line))
    if plan is not None:
        plan.close()
    if world > 1:
        barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

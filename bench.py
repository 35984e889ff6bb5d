#!/usr/bin/env python3
"""Headline benchmark: 3D energy-condition evaluation, fp64, n = 1024^3 (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1024] [--impl ours|reference]

One *step* = one full evaluation of the hot path on the synthetic (analytic) grid rho=5 sigma=4
v=1 extent=66, shells r<=40 / r<=60: fused kernel -> fixed-order tile reduction -> (N>1: NCCL
all-gather of the 245 KB chain sums) -> fixed-order row reduction -> 240-byte result to the host.

N = 1 runs in this process; N > 1 is launched by torchrun, one rank per GPU, each rank owning an
interleaved shard of the (j, k) tiles along the whole of axis 0 (weak scaling is NOT used: the volume
is fixed, so "scaling": "strong").

Printed JSON (one line, rank 0):
  value   : Gpoints/s, whole job, inputs (coordinates) resident in HBM -- CUDA events on the
            launching stream, summed over K steps, max over ranks; on one GPU the result row of every
            step stays on the device and all K are read (and checked against the warm-up's) after the
            closing barrier, on several GPUs every step ends with the blocking read of its result
  e2e     : same metric through the public API (backend.compute_energy_3d / distributed.
            compute_energy_3d_sharded) with host numpy coordinates copied from pinned memory and the
            result read back every step, wall clock
  roofline: FP64-pipe bound (no dense contraction, ~0 HBM bytes/point); see DESIGN.md §6
  cpu_baseline: the NumPy restatement of the reference (oracle/oracle_np.py) on every host core, bounded sample
`--impl reference` times that CPU path alone, on every host core (worker processes over slabs of the sample volume;
the reference itself is pure NumPy, single-threaded).
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
# FP64 issue slots per grid point.  CONTRACT = SURVEY.md §8(d) / BASELINE.md §4 (IEEE divisions at 9 slots
# each, exp/log1p on every point).  EXACT_ALG = the bit-exact algorithm this engine runs, each grid point
# evaluated once (no halo recompute), from the ncu per-instruction counts of the hot loops divided by their
# geometric redundancy: phi 35 (sqrt 8, two divisions 2 x 6, other 15), nine
# derivatives x 4, rho*dV 19, sums 5.  The kernel EXECUTES 113 FP64 instructions per point (halo recompute).
SLOTS_CONTRACT = 271.0
SLOTS_EXACT_ALG = 95.0
CPU_SAMPLE_N = 256
# dram__bytes_read.sum + dram__bytes_write.sum of k_energy3d, one launch at n = 1024^3 on one GPU, from the
# committed `ncu --set full` capture of the same kernel launch (profiles/r1_k_energy3d_n1024_session2_summary.txt)
# (111 616 B read, 0 B written in the capture of the final kernel; 490 496 B in the one before -- coordinate and
# instruction fetches, it varies with the state of the L2)
NCU_DRAM_BYTES_PER_LAUNCH = 111616 + 0


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


def _reference_slab(bounds):
    """Worker of the reference arm: the NumPy port on planes [a, b) of the sample volume (2-plane overlap recomputed)."""
    from oracle import oracle_np
    a, b = bounds
    r = oracle_np.energy_3d(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=CPU_SAMPLE_N, shells=SHELLS, i_begin=a, i_end=b)
    return r["e_pos"], r["e_neg"]


def _reference_timed(steps, warmup=1):
    """The NumPy port of the reference path on every host core: the sample volume (n = CPU_SAMPLE_N) is cut into slabs of
    axis 0 that worker processes evaluate independently, as the reference's own convergence_study_3d.py fans out over
    subprocesses (the reference itself is single-threaded NumPy).  Returns seconds per step and the worker count."""
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
    e_pos = sum(p[0] for p in parts)
    assert 1.02 < e_pos < 1.04, e_pos          # n = 256: 1.0274931410661474
    return dt, cores, avail, len(slabs)


def _reference_record(dt, cores, avail, nslabs, n_full):
    n = CPU_SAMPLE_N
    return {"value": n ** 3 / dt / 1e9, "unit": "Gpoints/s", "cores": cores, "kind": "port",
            "sample": f"each step = one n={n}^3 evaluation (bounded sample of the n={n_full}^3 workload: the reference needs "
                      f"~224 B/point, i.e. ~240 GB at 1024^3) as {nslabs} slabs of axis 0 on {cores} worker processes; "
                      f"NumPy port oracle/oracle_np.py (the reference itself is single-threaded)",
            "host_cores_available": avail}


def cpu_baseline(n_full=1024, reps=2):
    """cpu_baseline leg of the main arm: the same measurement as `--impl reference`, two steps."""
    return _reference_record(*_reference_timed(reps), n_full)


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host, every core (see _reference_timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(max(args.steps, 1), 5)
    dt, cores, avail, nslabs = _reference_timed(steps)
    rec = _reference_record(dt, cores, avail, nslabs, args.n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": "Gpoints/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D n={args.n}^3 fp64 energy-condition evaluation (BASELINE config 4): rho=5 sigma=4 v=1 "
                               f"extent=66, shells r<=40 and r<=60, exact (IEEE-replay) mode"},
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args(argv)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from irrotational_warp_lab_b200 import backend, distributed, sharding
    from irrotational_warp_lab_b200 import _capi
    from irrotational_warp_lab_b200._capi import ROW_WIDTH, TILE_CHAINS
    from irrotational_warp_lab_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
    tiles = sharding.tile_shard(world, rank) if world > 1 else None
    i_begin, i_end = (0, n) if (world == 1 or tiles is not None) else sharding.slab_bounds(n, world)[rank]
    nfb = sharding.num_flush_blocks(n)
    r0, r1 = sharding.rows_of_slab(i_begin, i_end)

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

    def resident_step():
        if plan is not None:
            plan.run_async(table.data_ptr(), stream.cuda_stream)
        if tiles is not None:
            full = sharding.all_gather_chains(table, world)
            return eng.reduce_chains(full.data_ptr(), world, nfb, len(SHELLS), stream.cuda_stream)
        full = sharding.all_gather_table(table[r0:r1], n, world) if world > 1 else table
        return eng.reduce_table(full.data_ptr(), nfb, len(SHELLS), stream.cuda_stream)

    # nvidia-smi is started BEFORE the warm-up and given time to initialise NVML: its start-up briefly stalls the driver,
    # which showed up as 0.2-0.3 ms outliers in the 1.4 ms steps of an 8-GPU run when it fell into the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    # Timed steps: the same launches with nothing read back in between -- the result row of every step stays on the
    # device (k_reduce_rows writes it to rows_out[s]) and all K rows are read after the closing synchronisation, so no host
    # thread sits inside the per-step event pairs (a late one used to show up as a 1-2 ms outlier in one step of ten).
    rows_out = torch.zeros((steps, ROW_WIDTH), dtype=torch.float64, device=dev)
    scratch_tab = torch.empty((nfb, ROW_WIDTH), dtype=torch.float64, device=dev)
    keep_alive = []

    def resident_step_async(s):
        if plan is not None:
            plan.run_async(table.data_ptr(), stream.cuda_stream)
        if tiles is not None:
            full = sharding.all_gather_chains(table, world)
            keep_alive.append(full)
            eng.reduce_chains_async(full.data_ptr(), world, nfb, scratch_tab.data_ptr(), rows_out[s].data_ptr(), stream.cuda_stream)
        else:
            full = sharding.all_gather_table(table[r0:r1], n, world) if world > 1 else table
            keep_alive.append(full)
            eng.reduce_table_async(full.data_ptr(), nfb, rows_out[s].data_ptr(), stream.cuda_stream)

    for _ in range(warmup):
        res = resident_step()                   # blocking flavour: also the value the timed steps must reproduce
    gc.collect()
    gc.disable()          # no collector pauses between the launches of the timed steps (one late rank delays every rank)
    barrier()
    t_region0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    # (the non-blocking flavour is used on one GPU, where it was validated on hardware; multi-GPU steps keep the blocking
    # read of the result, their outliers came from the clock sampler's start-up and are gone)
    use_async = (world == 1)
    for s in range(steps):
        flush_buf.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        if use_async:
            resident_step_async(s)
        else:
            res = resident_step()
        ev[s][1].record(stream)
    barrier()
    t_region1 = time.perf_counter()
    if use_async:
        rows_host = rows_out.cpu().numpy()
        for s in range(steps):
            got = sharding.reduce_table_host(rows_host[s:s + 1], len(SHELLS))
            assert (got["e_pos"], got["e_neg"], got["cnt_neg"], got["cnt_pos"]) == \
                   (res["e_pos"], res["e_neg"], res["cnt_neg"], res["cnt_pos"]), f"timed step {s} disagrees with the warm-up"
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
    # k_energy3d, k_reduce_tiles | k_chain_sums, (k_combine_chains,) k_reduce_rows
    launches_value = steps * ((2 if plan is not None else 0) + (2 if tiles is not None else 1))

    # kernel-only duration of the dominant kernel (rank 0's slab), CUDA events on its launching stream
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

    # ---------------- e2e: public API, host buffers, wall clock ------------------------------------------
    def api_step():
        if world > 1:
            return distributed.compute_energy_3d_sharded(rho=RHO, sigma=SIGMA, v=V, extent=EXTENT, n=n, shells=SHELLS)
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
    launches_e2e = steps * (4 if tiles is not None else 3)
    gc.enable()
    if rank == 0:
        sampler.stop()
    clocks = sampler.summary(t_region0, time.perf_counter()) if rank == 0 else None

    # sanity: the two paths agree bit-for-bit and look like the reference's convergence series
    assert run["energies"]["e_pos"] == res["e_pos"] and run["energies"]["e_neg"] == res["e_neg"], "paths disagree"
    if n == 1024:
        assert 1.058 < res["e_pos"] < 1.08 and -1.03 < res["e_neg"] < -1.005, res

    if rank == 0:
        peak = eng.measure_fma_peak(False)
        kernel_points = (i_end - i_begin) * n * n // (world if tiles is not None else 1)     # this rank's share
        pts_per_s_kernel = kernel_points / (kern_ms_avg * 1e-3) if kern_ms_avg else 0.0
        achieved_contract = pts_per_s_kernel * SLOTS_CONTRACT * 2 / 1e12
        achieved_exact = pts_per_s_kernel * SLOTS_EXACT_ALG * 2 / 1e12
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        # coordinates in; per-(flush block, tile) partial rows out (36 x 60 tiles of the default 40 x 64 geometry, 62 columns
        # in the tiles at the k faces);
        # they are consumed by k_reduce_tiles and usually never leave L2
        ntiles = _capi.load().iwb_debug_tiles_j(n, None, 0) * _capi.load().iwb_debug_tiles_k(n, None, 0)
        alg_bytes = 8.0 * (2 * n + n + n) + 8.0 * ROW_WIDTH * (r1 - r0) * ntiles / (world if tiles is not None else 1)
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D n={n}^3 fp64 energy-condition evaluation (BASELINE config 4): rho=5 sigma=4 v=1 "
                                   f"extent=66, shells r<=40 and r<=60, exact (IEEE-replay) mode",
                       "sharding": ("single GPU" if world == 1 else
                                    f"{world} interleaved tile shards (every {world}-th 36x60 tile of the (j,k) plane, whole axis 0), "
                                    f"all-gather of {nfb}x{TILE_CHAINS}x{ROW_WIDTH} f64 chain sums" if tiles is not None else
                                    f"{world} slab(s) of axis 0, analytic halo recompute, all-gather of a {nfb}x{ROW_WIDTH} f64 table"),
                       "l2": "256 MiB buffer written between timed steps (inputs are 32 KB of coordinates)",
                       "timing": "per-step CUDA-event pairs on the launching stream, summed; max over ranks"},
            "e2e": {"value": e2e_value, "unit": "Gpoints/s", "h2d_bytes_per_step": 8 * 4 * n,
                    "d2h_bytes_per_step": 8 * ROW_WIDTH, "ms_per_step": e2e_s / steps * 1e3,
                    "api": "distributed.compute_energy_3d_sharded" if world > 1 else "backend.compute_energy_3d"},
            "gpu_launches": launches_value + launches_e2e,
            "roofline": {
                "bound": "fp64", "kernel": "k_energy3d", "unit": "TFLOP/s",
                "achieved": achieved_exact, "peak": peak["tflops"], "frac": achieved_exact / peak["tflops"],
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (n == 1024 and world == 1) else None,
                "slots_per_point": SLOTS_EXACT_ALG,
                "note": "FP64-pipe bound: achieved = kernel points/s x 95 FP64 issue slots/point x 2 (slots of the "
                        "bit-exact algorithm with every point evaluated once, from ncu SASS counts, DESIGN.md §6; the "
                        "kernel executes 113/point including the phi halo recompute); peak = register-resident DFMA "
                        "microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 vector figure)",
                "kernel_ms": kern_ms_avg, "kernel_points": kernel_points,
                "contract": {"slots_per_point": SLOTS_CONTRACT, "achieved": achieved_contract,
                             "frac": achieved_contract / peak["tflops"],
                             "note": "SURVEY.md §8(d) figure (9-slot IEEE divisions, exp/log1p at every point)"},
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms_avg * 1e-3) / 1e9
                        if kern_ms_avg else None, "peak_gbs": hbm_peak, "note": "never binding: ~0 B/point"},
                "peak_sm_mhz_effective": peak["sm_mhz_effective"],
            },
            "clocks": clocks,
            "result": {"e_pos": res["e_pos"], "e_neg": res["e_neg"], "cnt_neg": res["cnt_neg"], "cnt_pos": res["cnt_pos"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
        print(json.dumps(line))
    if plan is not None:
        plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
